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
#include "gemm_kernel.cuh"

namespace {

// production configuration (chosen with tools/gemm_bench.cu on a B200)
using GramCfg = kfg::Cfg<KF_BK, KF_STAGES, 2, 4, false>;
static_assert(GramCfg::THREADS == KF_GEMM_THREADS, "thread count");
constexpr size_t GEMM_SMEM_BYTES = GramCfg::SMEM;

template <bool WEIGHTED>
__device__ __forceinline__ void gemm_tile_body(const KfGemmTask& t, double* smem) {
    kfg::gemm_tile_body<GramCfg, WEIGHTED, true>(t, smem);
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
