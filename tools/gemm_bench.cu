// Standalone benchmark of the DMMA tile kernel variants (gemm_kernel.cuh) on the config-5 task
// list: 10 Kronecker blocks x (36 G + 64 C) tiles over an L2-resident panel.  Prints TFLOP/s
// (flops issued) per variant and a checksum to confirm the variants agree.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../koopman-realizations_b200/csrc/gemm_kernel.cuh"

template <class C, bool W, bool PF>
__global__ void __launch_bounds__(C::THREADS, 1) bench_kernel(const KfGemmTask* __restrict__ tasks) {
    extern __shared__ __align__(16) double smem[];
    const KfGemmTask t = tasks[blockIdx.x];
    if (W && t.W == nullptr) { kfg::gemm_tile_body<C, false, PF>(t, smem); return; }
    kfg::gemm_tile_body<C, W, PF>(t, smem);
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

template <class C, bool W, bool PF>
void run(const char* name, const KfGemmTask* d_tasks, int ntasks, int Mc, double* accum, size_t accum_n, double* d_sum) {
    cudaFuncSetAttribute(bench_kernel<C, W, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    cudaMemset(accum, 0, accum_n * 8);
    bench_kernel<C, W, PF><<<ntasks, C::THREADS, C::SMEM>>>(d_tasks);
    cudaMemset(d_sum, 0, 8);
    checksum<<<1, 1024>>>(accum, accum_n, d_sum);
    double h = 0; cudaMemcpy(&h, d_sum, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("%-34s ERROR %s\n", name, cudaGetErrorString(e)); return; }
    bench_kernel<C, W, PF><<<ntasks, C::THREADS, C::SMEM>>>(d_tasks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 10;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) bench_kernel<C, W, PF><<<ntasks, C::THREADS, C::SMEM>>>(d_tasks);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    double tf = (double)ntasks * 2.0 * 128 * 128 * Mc / ms / 1e9;
    printf("%-34s thr=%4d smem=%6zu  %.3f ms  %.2f TF  checksum %.10e\n", name, C::THREADS, C::SMEM, ms, tf, h);
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
    for (int q = 0; q < nW; ++q) {
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
    }
    cudaMalloc(&accum, ntile * 16384 * 8);
    for (auto& g : tasks) g.out = accum + ((size_t)g.out / 8);
    KfGemmTask* d_tasks; cudaMalloc(&d_tasks, tasks.size() * sizeof(KfGemmTask));
    cudaMemcpy(d_tasks, tasks.data(), tasks.size() * sizeof(KfGemmTask), cudaMemcpyHostToDevice);
    cudaMalloc(&d_sum, 8);
    printf("tasks=%zu Mc=%d panel=%.1f MB accum=%.1f MB\n", tasks.size(), Mc, rows * (double)Mc * 8 / 1e6, ntile * 16384 * 8 / 1e6);
    const int nt = (int)tasks.size();
    using namespace kfg;
    run<Cfg<16, 4, 2, 4, false>, true, false>("BK16 S4 2x4 lds64 (r1 baseline)", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    run<Cfg<16, 4, 2, 4, false>, true, true>("BK16 S4 2x4 lds64 +prefetch", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    run<Cfg<16, 4, 2, 4, false>, false, false>("BK16 S4 2x4 lds64 unweighted", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    run<Cfg<32, 3, 2, 4, false>, true, true>("BK32 S3 2x4 lds64 +pf", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    run<Cfg<16, 4, 2, 4, true>, true, true>("BK16 S4 2x4 lds128 +pf", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    run<Cfg<32, 2, 2, 4, true>, true, true>("BK32 S2 2x4 lds128 +pf", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    run<Cfg<16, 5, 2, 4, false>, true, true>("BK16 S5 2x4 lds64 +pf", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    run<Cfg<16, 4, 4, 4, false>, true, true>("BK16 S4 4x4 lds64 +pf (16 warps)", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    run<Cfg<16, 4, 4, 4, true>, true, true>("BK16 S4 4x4 lds128 +pf (16 warps)", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    run<Cfg<32, 3, 4, 4, false>, true, true>("BK32 S3 4x4 lds64 +pf (16 warps)", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    run<Cfg<16, 4, 4, 2, false>, true, true>("BK16 S4 4x2 lds64 +pf (32x64 wt)", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    run<Cfg<16, 4, 2, 8, false>, true, true>("BK16 S4 2x8 lds64 +pf (64x16 wt)", d_tasks, nt, Mc, accum, ntile * 16384, d_sum);
    return 0;
}
