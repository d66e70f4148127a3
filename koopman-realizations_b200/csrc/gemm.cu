// FP64 tensor-core (DMMA) contraction kernel for sm_100a.
//
//   out[m][n] (+)= alpha * sum_k w[k] * A[m][k] * B[n][k]
//
// This is the Psi'Psi / Psi'Y contraction of the fit (Ksysid.m:1114, 1125) over one
// L2-resident lifted panel (rows = observables, k = snapshots of the chunk), and the
// same kernel serves the solver's trailing updates.  On sm_100a the FP64 tensor path is
// the warp-level mma.sync m8n8k4 (SASS DMMA.8x8x4); tcgen05 has no f64 kind.
//
// CTA: 128x128 output tile, 256 threads = 8 warps (2 x 4), warp tile 64x32 =
// 8x4 DMMA tiles -> 64 accumulator doubles per thread.  Operands stream global/L2 ->
// shared through a 4-stage cp.async (LDGSTS) pipeline, 16 contraction elements per stage.
// Shared rows are padded to 20 doubles so the per-fragment LDS.64 (row = lane/4, k = lane%4)
// hits 16 distinct 8-byte bank pairs per half-warp: conflict-free without swizzling.
#include "kf_internal.h"

namespace {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

constexpr int A_STAGE = KF_BM * KF_LDS;   // doubles per A stage
constexpr int B_STAGE = KF_BN * KF_LDS;
constexpr size_t GEMM_SMEM_BYTES = (size_t)KF_STAGES * (A_STAGE + B_STAGE + KF_BK) * sizeof(double);

template <bool WEIGHTED>
__device__ __forceinline__ void gemm_tile_body(const KfGemmTask& t, double* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    double* As = smem;
    double* Bs = smem + KF_STAGES * A_STAGE;
    double* Ws = Bs + KF_STAGES * B_STAGE;

    // global->shared assignment: 128 rows x 8 sixteen-byte chunks per operand per stage
    const int lrow = tid >> 3, lkc = (tid & 7) * 2;
    const double* a_src[4];
    const double* b_src[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = lrow + i * 32;
        a_src[i] = t.A + (long long)min(row, t.a_rows - 1) * t.lda + lkc;
        b_src[i] = t.B + (long long)min(row, t.b_rows - 1) * t.ldb + lkc;
    }
    const int nk = (t.k1 - t.k0) / KF_BK;

    auto load_stage = [&](int stage, int kt) {
        const int k = t.k0 + kt * KF_BK;
        double* as = As + stage * A_STAGE + lrow * KF_LDS + lkc;
        double* bs = Bs + stage * B_STAGE + lrow * KF_LDS + lkc;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cp_async16(as + i * 32 * KF_LDS, a_src[i] + k);
            cp_async16(bs + i * 32 * KF_LDS, b_src[i] + k);
        }
        if (WEIGHTED && tid < 8) cp_async16(Ws + stage * KF_BK + tid * 2, t.W + k + tid * 2);
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < KF_STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    const int frag = (lane >> 2) * KF_LDS + (lane & 3);
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<KF_STAGES - 2>();
        __syncthreads();
        {
            const int nxt = kt + KF_STAGES - 1;
            if (nxt < nk) load_stage(nxt % KF_STAGES, nxt);
            cp_async_commit();
        }
        const int stage = kt % KF_STAGES;
        const double* as = As + stage * A_STAGE + (wm * 64) * KF_LDS + frag;
        const double* bs = Bs + stage * B_STAGE + (wn * 32) * KF_LDS + frag;
        const double* ws = Ws + stage * KF_BK + (lane & 3);
#pragma unroll
        for (int kk = 0; kk < KF_BK / 4; ++kk) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = as[i * 8 * KF_LDS + kk * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = bs[j * 8 * KF_LDS + kk * 4];
            if (WEIGHTED) {
                const double w = ws[kk * 4];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] *= w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds (m = lane/4, n = 2*(lane%4)+{0,1}) of every 8x8 DMMA tile
    const int m_base = wm * 64 + (lane >> 2);
    const int n_base = wn * 32 + 2 * (lane & 3);
    const bool vec = (t.ldn == 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m_base + i * 8;
        if (m >= t.a_rows) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n_base + j * 8;
            double v0 = t.alpha * acc[i][j][0], v1 = t.alpha * acc[i][j][1];
            double* p = t.out + (long long)m * t.ldm + (long long)n * t.ldn;
            if (vec && n + 1 < t.b_rows) {
                double2* p2 = reinterpret_cast<double2*>(p);
                if (t.accumulate) {
                    double2 o = *p2;
                    v0 += o.x;
                    v1 += o.y;
                }
                *p2 = make_double2(v0, v1);
            } else {
                if (n < t.b_rows) {
                    if (t.accumulate) v0 += p[0];
                    p[0] = v0;
                }
                if (n + 1 < t.b_rows) {
                    if (t.accumulate) v1 += p[t.ldn];
                    p[t.ldn] = v1;
                }
            }
        }
    }
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(KF_GEMM_THREADS, 1) kf_gram_tile_kernel(const KfGemmTask* __restrict__ tasks) {
    extern __shared__ __align__(16) double kf_smem[];
    const KfGemmTask t = tasks[blockIdx.x];
    if (WEIGHTED) {
        if (t.W == nullptr) {   // mixed task lists: unweighted tasks run the plain body
            gemm_tile_body<false>(t, kf_smem);
            return;
        }
    }
    gemm_tile_body<WEIGHTED>(t, kf_smem);
}

__global__ void __launch_bounds__(KF_GEMM_THREADS, 1) kf_gemm_grid_kernel(const KfGemmGrid g) {
    extern __shared__ __align__(16) double kf_smem[];
    const int tm = blockIdx.y, tn = blockIdx.x;
    if (g.lower_only && tn > tm) return;
    KfGemmTask t;
    t.A = g.A + (long long)tm * KF_BM * g.lda;
    t.B = g.B + (long long)tn * KF_BN * g.ldb;
    t.W = nullptr;
    t.out = g.out + (long long)tm * KF_BM * g.ldm + (long long)tn * KF_BN * g.ldn;
    t.lda = g.lda; t.ldb = g.ldb; t.ldm = g.ldm; t.ldn = g.ldn;
    t.k0 = g.k0; t.k1 = g.k1;
    t.a_rows = min(KF_BM, g.m - tm * KF_BM);
    t.b_rows = min(KF_BN, g.n - tn * KF_BN);
    t.alpha = g.alpha;
    t.accumulate = g.accumulate;
    gemm_tile_body<false>(t, kf_smem);
}

bool g_attr_set = false;
cudaError_t ensure_attrs() {
    if (g_attr_set) return cudaSuccess;
    cudaError_t e;
    e = cudaFuncSetAttribute(kf_gram_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kf_gram_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kf_gemm_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    g_attr_set = true;
    return cudaSuccess;
}

}  // namespace

int kf_launch_gemm_tasks(kf_ctx* ctx, const KfGemmTask* d_tasks, int ntasks, bool weighted, cudaStream_t st) {
    if (ntasks <= 0) return KF_OK;
    KF_CUDA(ctx, ensure_attrs());
    if (weighted)
        kf_gram_tile_kernel<true><<<ntasks, KF_GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(d_tasks);
    else
        kf_gram_tile_kernel<false><<<ntasks, KF_GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(d_tasks);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}

int kf_launch_gemm_grid(kf_ctx* ctx, const KfGemmGrid& g, cudaStream_t st) {
    if (g.m <= 0 || g.n <= 0 || g.k1 <= g.k0) return KF_OK;
    KF_CUDA(ctx, ensure_attrs());
    dim3 grid((g.n + KF_BN - 1) / KF_BN, (g.m + KF_BM - 1) / KF_BM);
    kf_gemm_grid_kernel<<<grid, KF_GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(g);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    double tiles = (double)grid.x * grid.y;
    if (g.lower_only) tiles = (double)grid.y * (grid.y + 1) / 2;
    ctx->dmma_flops += tiles * 2.0 * KF_BM * KF_BN * (double)(g.k1 - g.k0);
    return KF_OK;
}
