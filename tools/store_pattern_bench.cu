// Which store pattern reaches the HBM write rate for a feature-major matrix out[row * ld + snapshot] that is produced tile by
// tile (a tile = LS consecutive snapshots x all rows)?  Decides the design of the materialising lift (lift.cu):
//   stg   : 128-bit st.global per thread, a warp covers 32 * 16 B = 512 B = (512 / (8 LS)) rows x one LS-snapshot run
//   tma   : cp.async.bulk.tensor.2d shared -> global (UTMASTG), box = LS snapshots x RB rows, one elected thread
//   bulk  : cp.async.bulk.global.shared::cta, one 8 LS-byte run per row
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern_bench tools/store_pattern_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int LS>
__global__ void __launch_bounds__(256) stg_kernel(double* out, long long ld, int R, long long ntiles, int hint) {
    constexpr int PS = LS / 2, NR = 256 / PS;
    const int p2 = (threadIdx.x % PS) * 2, rr = threadIdx.x / PS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        double* base = out + tile * LS + p2;
        const double2 v = make_double2((double)tile, (double)p2);
        for (int r = rr; r < R; r += NR) {
            double2* dst = reinterpret_cast<double2*>(base + (long long)r * ld);
            if (hint == 1) __stcs(dst, v);
            else if (hint == 2) __stwt(dst, v);
            else *dst = v;
        }
    }
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// box = LS x RB doubles in shared memory, filled once; every (tile, row box) is one tensor store
__global__ void __launch_bounds__(128) tma_kernel(const __grid_constant__ CUtensorMap tm, int LS, int RB, int R, long long ntiles, int depth) {
    extern __shared__ __align__(128) double sm[];
    for (int i = threadIdx.x; i < LS * RB; i += blockDim.x) sm[i] = (double)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nrb = (R + RB - 1) / RB;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int b = 0; b < nrb; ++b) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                             :: "l"(&tm), "r"((int)(tile * LS)), "r"(b * RB), "r"(smem_u32(sm)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else if (depth == 4) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 16;" ::: "memory");
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// one bulk copy of 8 LS bytes per row, issued by all lanes of the CTA
__global__ void __launch_bounds__(128) bulk_kernel(double* out, long long ld, int LS, int R, long long ntiles) {
    extern __shared__ __align__(128) double sm[];
    for (int i = threadIdx.x; i < LS * 64; i += blockDim.x) sm[i] = (double)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int r = threadIdx.x; r < R; r += blockDim.x) {
            double* dst = out + (long long)r * ld + tile * LS;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(dst), "r"(smem_u32(sm + (r & 63) * LS)), "r"(LS * 8) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <class F>
double time_ms(F f, int reps = 5) {
    f(); f();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    encode_fn enc = nullptr;
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        enc = (encode_fn)p;
    }
    const int R = 8192;
    double* out;
    const long long Mmax = 133120;
    CK(cudaMalloc(&out, (size_t)R * Mmax * 8));
    for (long long M : {133000LL, 133120LL}) {
        const double gb = (double)R * M * 8 / 1e9;
        printf("== R = %d rows, M = ld = %lld (row stride %s128-byte aligned), %.2f GB per pass\n", R, M, (M * 8) % 128 ? "NOT " : "", gb);
        for (int hint = 0; hint < 3; ++hint)
            for (int per_sm : {2, 4, 8}) {
                const int g = 148 * per_sm;
#define RUN(LSV) { const long long nt = M / LSV; double ms = time_ms([&] { stg_kernel<LSV><<<g, 256>>>(out, M, R, nt, hint); }); \
                   printf("stg  LS %3d hint %d ctas/sm %d : %.3f ms  %.0f GB/s\n", LSV, hint, per_sm, ms, (double)R * nt * LSV * 8 / ms / 1e6); }
                RUN(16) RUN(32) RUN(64) RUN(128) RUN(256)
            }
        for (int LS : {16, 32, 64, 128})
            for (int RB : {64, 256}) {
                if ((size_t)LS * RB * 8 > 200 * 1024) continue;
                CUtensorMap tm;
                const cuuint64_t gdim[2] = {(cuuint64_t)M, (cuuint64_t)R};
                const cuuint64_t gstr[1] = {(cuuint64_t)M * 8};
                const cuuint32_t box[2] = {(cuuint32_t)LS, (cuuint32_t)RB};
                const cuuint32_t es[2] = {1, 1};
                CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, out, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("tma LS %d RB %d: encode failed %d\n", LS, RB, (int)r); continue; }
                const size_t smem = (size_t)LS * RB * 8;
                CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                for (int per_sm : {1, 2, 4})
                    for (int depth : {1, 4, 16}) {
                        if (smem * per_sm > 220 * 1024) continue;
                        const long long nt = M / LS;
                        double ms = time_ms([&] { tma_kernel<<<148 * per_sm, 128, smem>>>(tm, LS, RB, R, nt, depth); });
                        printf("tma  LS %3d RB %3d ctas/sm %d depth %2d : %.3f ms  %.0f GB/s\n", LS, RB, per_sm, depth, ms, (double)R * nt * LS * 8 / ms / 1e6);
                    }
            }
        CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int LS : {16, 64, 256})
            for (int per_sm : {2, 4}) {
                const long long nt = M / LS;
                double ms = time_ms([&] { bulk_kernel<<<148 * per_sm, 128, (size_t)LS * 64 * 8>>>(out, M, LS, R, nt); });
                printf("bulk LS %3d ctas/sm %d : %.3f ms  %.0f GB/s\n", LS, per_sm, ms, (double)R * nt * LS * 8 / ms / 1e6);
            }
    }
    CK(cudaGetLastError());
    return 0;
}
