// Standalone benchmark of the DMMA tile kernel variants (gemm_kernel.cuh) on the config-5 task
// list: 10 Kronecker blocks x (36 G + 64 C) tiles over an L2-resident panel.  Prints TFLOP/s
// (flops issued) per variant and a checksum to confirm the variants agree; plus register-level
// probes of the warp-tile inner loop (no global memory) to separate pipe limits from feeding limits.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../koopman-realizations_b200/csrc/gemm_kernel.cuh"
#include "../koopman-realizations_b200/csrc/tma_host.h"

template <class C, bool W, bool PF, bool BULK = false>
__global__ void __launch_bounds__(C::THREADS, C::MINB) bench_kernel(const KfGemmTask* __restrict__ tasks) {
    extern __shared__ __align__(16) double smem[];
    const KfGemmTask t = tasks[blockIdx.x];
    if constexpr (BULK) {
        if (W && t.W == nullptr) { kfg::gemm_tile_body_bulk<C, false, PF>(t, smem); return; }
        kfg::gemm_tile_body_bulk<C, W, PF>(t, smem);
    } else {
        if (W && t.W == nullptr) { kfg::gemm_tile_body<C, false, PF>(t, smem); return; }
        kfg::gemm_tile_body<C, W, PF>(t, smem);
    }
}

__global__ void fill(double* p, size_t n, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned x = (unsigned)(i * 2654435761u) ^ seed; x ^= x >> 13; x *= 0x5bd1e995u; x ^= x >> 15;
        p[i] = (double)(x & 0xffff) / 65536.0 - 0.5;
    }
}
__global__ void checksum(const double* p, size_t n, double* out) {
    double s = 0; for (size_t i = threadIdx.x; i < n; i += blockDim.x) s += p[i] * (double)((i % 7) + 1);
    atomicAdd(out, s);
}

template <class C, bool W, bool PF, bool BULK = false>
void run(const char* name, const std::vector<KfGemmTask>& base, int Mc, double* accum, size_t accum_n, double* d_sum) {
    // split every 128x128 tile into (128/BM) x (128/BN) CTA tasks
    std::vector<KfGemmTask> tasks;
    for (const auto& b : base)
        for (int sm = 0; sm < 128 / C::BM; ++sm)
            for (int sn = 0; sn < 128 / C::BN; ++sn) {
                KfGemmTask g = b;
                g.A = b.A + (size_t)sm * C::BM * b.lda;
                g.B = b.B + (size_t)sn * C::BN * b.ldb;
                g.out = b.out + (size_t)sm * C::BM * 128 + sn * C::BN;
                g.a_rows = C::BM; g.b_rows = C::BN;
                tasks.push_back(g);
            }
    const int ntasks = (int)tasks.size();
    KfGemmTask* d_tasks; cudaMalloc(&d_tasks, tasks.size() * sizeof(KfGemmTask));
    cudaMemcpy(d_tasks, tasks.data(), tasks.size() * sizeof(KfGemmTask), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(bench_kernel<C, W, PF, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BULK ? kfg::bulk_smem_bytes<C>() : C::SMEM));
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bench_kernel<C, W, PF, BULK>, C::THREADS, (BULK ? kfg::bulk_smem_bytes<C>() : C::SMEM));
    cudaMemset(accum, 0, accum_n * 8);
    bench_kernel<C, W, PF, BULK><<<ntasks, C::THREADS, (BULK ? kfg::bulk_smem_bytes<C>() : C::SMEM)>>>(d_tasks);
    cudaMemset(d_sum, 0, 8);
    checksum<<<1, 1024>>>(accum, accum_n, d_sum);
    double h = 0; cudaMemcpy(&h, d_sum, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("%-40s ERROR %s\n", name, cudaGetErrorString(e)); cudaFree(d_tasks); return; }
    bench_kernel<C, W, PF, BULK><<<ntasks, C::THREADS, (BULK ? kfg::bulk_smem_bytes<C>() : C::SMEM)>>>(d_tasks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 10;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) bench_kernel<C, W, PF, BULK><<<ntasks, C::THREADS, (BULK ? kfg::bulk_smem_bytes<C>() : C::SMEM)>>>(d_tasks);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    double tf = (double)base.size() * 2.0 * 128 * 128 * Mc / ms / 1e9;
    printf("%-40s thr=%4d smem=%6zu occ=%d ctas=%5d  %.3f ms  %.2f TF  checksum %.10e\n", name, C::THREADS, (BULK ? kfg::bulk_smem_bytes<C>() : C::SMEM), occ, ntasks, ms, tf, h);
    cudaFree(d_tasks);
}

template <class C, bool W>
__global__ void __launch_bounds__(C::THREADS, C::MINB) tma_kernel(const KfTmaTask* __restrict__ tasks, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const KfTmaTask t = tasks[blockIdx.x];
    if (W && t.W == nullptr) { kfg::gemm_tile_body_tma<C, false>(t, &tmap, smem_raw); return; }
    kfg::gemm_tile_body_tma<C, W>(t, &tmap, smem_raw);
}

template <class C>
void run_tma(const char* name, const std::vector<KfGemmTask>& base, double* panel, int rows, int Mc, double* accum, size_t accum_n, double* d_sum) {
    CUtensorMap tmap;
    int rc = kf_make_panel_tensor_map(&tmap, panel, Mc, rows);
    if (rc) { printf("%-40s tensor map encode failed rc=%d\n", name, rc); return; }
    std::vector<KfTmaTask> tasks;
    for (const auto& b : base)
        for (int sn = 0; sn < 2; ++sn) {
            KfTmaTask g{};
            g.a_row = (int)((b.A - panel) / Mc);
            g.b_row = (int)((b.B - panel) / Mc) + sn * 64;
            g.W = b.W; g.w_row = b.W ? 0 : -1;
            g.k0 = 0; g.k1 = Mc;
            g.out = b.out + sn * 64;
            tasks.push_back(g);
        }
    const int ntasks = (int)tasks.size();
    KfTmaTask* d_tasks; cudaMalloc(&d_tasks, tasks.size() * sizeof(KfTmaTask));
    cudaMemcpy(d_tasks, tasks.data(), tasks.size() * sizeof(KfTmaTask), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(tma_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tma_kernel<C, true>, C::THREADS, C::SMEM);
    cudaMemset(accum, 0, accum_n * 8);
    tma_kernel<C, true><<<ntasks, C::THREADS, C::SMEM>>>(d_tasks, tmap);
    cudaMemset(d_sum, 0, 8);
    checksum<<<1, 1024>>>(accum, accum_n, d_sum);
    double h = 0; cudaMemcpy(&h, d_sum, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("%-40s ERROR %s\n", name, cudaGetErrorString(e)); cudaFree(d_tasks); return; }
    tma_kernel<C, true><<<ntasks, C::THREADS, C::SMEM>>>(d_tasks, tmap);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 10;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) tma_kernel<C, true><<<ntasks, C::THREADS, C::SMEM>>>(d_tasks, tmap);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    double tf = (double)base.size() * 2.0 * 128 * 128 * Mc / ms / 1e9;
    printf("%-40s thr=%4d smem=%6zu occ=%d ctas=%5d  %.3f ms  %.2f TF  checksum %.10e\n", name, C::THREADS, (size_t)C::SMEM, occ, ntasks, ms, tf, h);
    cudaFree(d_tasks);
}

// ---- register-level probes: the warp-tile inner loop without global memory
template <int MI, int NJ, bool LDS_, bool SYNC, bool WEIGHT>
__global__ void __launch_bounds__(256, 1) probe_kernel(double* out, int iters) {
    __shared__ double sm[2 * 128 * 20 + 16];
    for (int i = threadIdx.x; i < 2 * 128 * 20 + 16; i += 256) sm[i] = 1e-3 * (i % 13);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wm = warp >> 2, wn = warp & 3;
    double acc[MI][NJ][2];
    for (int i = 0; i < MI; ++i) for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0;
    const double* as = sm + (wm * 8 * MI + (lane >> 2)) * 20 + (lane & 3);
    const double* bs = sm + 128 * 20 + (wn * 8 * NJ + (lane >> 2)) * 20 + (lane & 3);
    double a[MI], b[NJ];
    for (int i = 0; i < MI; ++i) a[i] = as[i * 8 * 20];
    for (int j = 0; j < NJ; ++j) b[j] = bs[j * 8 * 20];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            if (LDS_) {
#pragma unroll
                for (int i = 0; i < MI; ++i) a[i] = as[i * 8 * 20 + kk * 4];
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[j] = bs[j * 8 * 20 + kk * 4];
            }
            if (WEIGHT) {
                const double w = sm[2 * 128 * 20 + kk * 4 + (lane & 3)];
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[j] *= w;
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) kfg::dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        if (SYNC) __syncthreads();
    }
    double s = 0;
    for (int i = 0; i < MI; ++i) for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int MI, int NJ, bool LDS_, bool SYNC, bool WEIGHT>
void probe(const char* name, double* out) {
    const int iters = 2000;
    probe_kernel<MI, NJ, LDS_, SYNC, WEIGHT><<<148, 256>>>(out, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe_kernel<MI, NJ, LDS_, SYNC, WEIGHT><<<148, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double tf = 148.0 * 8 * iters * 4.0 * MI * NJ * 512.0 / ms / 1e9;
    printf("probe %-44s %.3f ms  %.2f TF\n", name, ms, tf);
}

int main(int argc, char** argv) {
    const int N = 1024, Mc = argc > 1 ? atoi(argv[1]) : 1536, nW = 10;
    const int rows = 2 * N + 16;
    double *panel, *accum, *d_sum;
    const int tm = N / 128;
    std::vector<KfGemmTask> tasks;
    size_t ntile = 0;
    cudaMalloc(&panel, (size_t)rows * Mc * 8);
    fill<<<1024, 256>>>(panel, (size_t)rows * Mc, 17u);
    for (int q = 0; q < nW; ++q)
        for (int kind = 0; kind < 2; ++kind)
            for (int a = 0; a < tm; ++a)
                for (int b = 0; b < (kind == 0 ? a + 1 : tm); ++b) {
                    KfGemmTask g{};
                    g.A = panel + (size_t)(a * 128) * Mc;
                    g.B = panel + (size_t)((kind ? N : 0) + b * 128) * Mc;
                    g.W = q ? panel + (size_t)(2 * N + q) * Mc : nullptr;
                    g.out = (double*)(ntile * 16384 * 8);   // offset, fixed up below
                    g.lda = g.ldb = Mc; g.ldm = 128; g.ldn = 1; g.k0 = 0; g.k1 = Mc; g.a_rows = g.b_rows = 128;
                    g.alpha = 1.0; g.accumulate = 1;
                    tasks.push_back(g); ++ntile;
                }
    cudaMalloc(&accum, ntile * 16384 * 8);
    for (auto& g : tasks) g.out = accum + ((size_t)g.out / 8);
    cudaMalloc(&d_sum, 8);
    printf("tiles=%zu Mc=%d panel=%.1f MB accum=%.1f MB\n", tasks.size(), Mc, rows * (double)Mc * 8 / 1e6, ntile * 16384 * 8 / 1e6);
    using namespace kfg;
    const size_t an = ntile * 16384;
    probe<8, 4, false, false, false>("64x32 warp tile, regs only (no LDS, no sync)", accum);
    probe<8, 4, true, false, false>("64x32 + LDS fragments", accum);
    probe<8, 4, true, true, false>("64x32 + LDS + __syncthreads/16k", accum);
    probe<8, 4, true, true, true>("64x32 + LDS + sync + weights", accum);
    probe<4, 4, true, true, true>("32x32 + LDS + sync + weights", accum);
    probe<4, 2, true, true, true>("32x16 + LDS + sync + weights", accum);
    run<Cfg<16, 3, 2, 2, true, 128, 64, 2>, true, true>("128x64 BK16 S3 2x2 lds128 occ2 (prod)", tasks, Mc, accum, an, d_sum);
    run_tma<kfg::TmaCfg<3>>("TMA tensor-map swz128 128x64 S3 occ2", tasks, panel, rows, Mc, accum, an, d_sum);
    run_tma<kfg::TmaCfg<4>>("TMA tensor-map swz128 128x64 S4 occ2", tasks, panel, rows, Mc, accum, an, d_sum);
    run_tma<kfg::TmaCfg<6>>("TMA tensor-map swz128 128x64 S6 occ1?", tasks, panel, rows, Mc, accum, an, d_sum);
    return 0;
}
