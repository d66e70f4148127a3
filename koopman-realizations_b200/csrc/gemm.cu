// FP64 tensor-core (DMMA) contraction kernels for sm_100a.
//
//   out[m][n] (+)= alpha * sum_k w[k] * A[m][k] * B[n][k]
//
// This is the Psi'Psi / Psi'Y contraction of the fit (Ksysid.m:1114, 1125) over one
// L2-resident lifted panel (rows = observables, k = snapshots of the chunk), and the
// same kernel serves the solver's trailing updates.  The tile body lives in
// gemm_kernel.cuh; the configuration below was chosen by measurement on a B200
// (tools/gemm_bench.cu -> profiles/r01_gemm_variants_v2.txt):
//   CTA tile 128 x 64, 128 threads = 4 warps (2 x 2), warp tile 64 x 32 = 8 x 4 DMMA.8x8x4
//   tiles, 3-stage cp.async pipeline of 16 snapshots, LDS.128 fragments, 2 CTAs per SM so the
//   prologue / barrier / read-modify-write epilogue of one CTA hides under the other's DMMAs.
#include "gemm_kernel.cuh"

namespace {

using GramCfg = kfg::Cfg<KF_BK, KF_STAGES, 2, 2, true, KF_CTA_M, KF_CTA_N, 2>;
static_assert(GramCfg::THREADS == KF_GEMM_THREADS, "thread count");
constexpr size_t GEMM_SMEM_BYTES = GramCfg::SMEM;

template <bool WEIGHTED>
__global__ void __launch_bounds__(KF_GEMM_THREADS, GramCfg::MINB) kf_gram_tile_kernel(const KfGemmTask* __restrict__ tasks) {
    extern __shared__ __align__(16) double kf_smem[];
    const KfGemmTask t = tasks[blockIdx.x];
    if (WEIGHTED) {
        if (t.W == nullptr) {   // mixed task lists: unweighted tasks run the plain body
            kfg::gemm_tile_body<GramCfg, false, true>(t, kf_smem);
            return;
        }
    }
    kfg::gemm_tile_body<GramCfg, WEIGHTED, true>(t, kf_smem);
}

__global__ void __launch_bounds__(KF_GEMM_THREADS, GramCfg::MINB) kf_gemm_grid_kernel(const KfGemmGrid g) {
    extern __shared__ __align__(16) double kf_smem[];
    const int tm = blockIdx.y, tn = blockIdx.x;
    if (g.lower_only && tn * KF_CTA_N > tm * KF_CTA_M + KF_CTA_M - 1) return;
    KfGemmTask t;
    t.A = g.A + (long long)tm * KF_CTA_M * g.lda;
    t.B = g.B + (long long)tn * KF_CTA_N * g.ldb;
    t.W = nullptr;
    t.out = g.out + (long long)tm * KF_CTA_M * g.ldm + (long long)tn * KF_CTA_N * g.ldn;
    t.lda = g.lda; t.ldb = g.ldb; t.ldm = g.ldm; t.ldn = g.ldn;
    t.k0 = g.k0; t.k1 = g.k1;
    if (gridDim.z > 1) {      // split-K: chunk z of the k range (multiples of KF_BK) -> slab z
        const int kc = ((g.k1 - g.k0 + (int)gridDim.z - 1) / (int)gridDim.z + KF_BK - 1) / KF_BK * KF_BK;
        t.k0 = g.k0 + (int)blockIdx.z * kc;
        t.k1 = min(g.k1, t.k0 + kc);
        t.out += (long long)blockIdx.z * g.slab;
    }
    t.a_rows = min(KF_CTA_M, g.m - tm * KF_CTA_M);
    t.b_rows = min(KF_CTA_N, g.n - tn * KF_CTA_N);
    t.alpha = g.alpha;
    t.accumulate = g.accumulate;
    kfg::gemm_tile_body<GramCfg, false, false>(t, kf_smem);
}

// out[m][n] (+)= alpha * sum_k A[m][k] * B[k][n], B k-major (gemm_kernel.cuh body 1b); n padded to a multiple of 64
__global__ void __launch_bounds__(KF_GEMM_THREADS, GramCfg::MINB) kf_gemm_bkmajor_kernel(const KfGemmGrid g) {
    extern __shared__ __align__(16) double kf_smem[];
    const int tm = blockIdx.y, tn = blockIdx.x;
    KfGemmTask t;
    t.A = g.A + (long long)tm * KF_CTA_M * g.lda;
    t.B = g.B + (long long)tn * KF_CTA_N;
    t.W = nullptr;
    t.out = g.out + (long long)tm * KF_CTA_M * g.ldm + (long long)tn * KF_CTA_N * g.ldn;
    t.lda = g.lda; t.ldb = g.ldb; t.ldm = g.ldm; t.ldn = g.ldn;
    t.k0 = g.k0; t.k1 = g.k1;
    t.a_rows = min(KF_CTA_M, g.m - tm * KF_CTA_M);
    t.b_rows = KF_CTA_N;
    t.alpha = g.alpha;
    t.accumulate = g.accumulate;
    kfg::gemm_tile_body_bkmajor<GramCfg>(t, kf_smem);
}

// Gram kernel, TMA operand path (gemm_kernel.cuh body 3): tensor-map loads with SWIZZLE_128B + mbarrier ring.
using TmaGramCfg = kfg::TmaCfg<4>;

template <bool WEIGHTED>
__global__ void __launch_bounds__(TmaGramCfg::THREADS, TmaGramCfg::MINB)
kf_gram_tma_kernel(const KfTmaTask* __restrict__ tasks, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char kf_smem_raw[];
    const KfTmaTask t = tasks[blockIdx.x];
    if (WEIGHTED) {
        if (t.W == nullptr) {
            kfg::gemm_tile_body_tma<TmaGramCfg, false>(t, &tmap, kf_smem_raw);
            return;
        }
    }
    kfg::gemm_tile_body_tma<TmaGramCfg, WEIGHTED>(t, &tmap, kf_smem_raw);
}

cudaError_t ensure_attrs(kf_ctx* ctx) {
    cudaError_t e;
    if ((e = kf_ensure_smem(ctx, kf_gram_tile_kernel<false>, (size_t)GEMM_SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = kf_ensure_smem(ctx, kf_gram_tile_kernel<true>, (size_t)GEMM_SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = kf_ensure_smem(ctx, kf_gemm_grid_kernel, (size_t)GEMM_SMEM_BYTES)) != cudaSuccess) return e;
    if ((e = kf_ensure_smem(ctx, kf_gemm_bkmajor_kernel, kfg::bkmajor_smem_bytes<GramCfg>())) != cudaSuccess) return e;
    if ((e = kf_ensure_smem(ctx, kf_gram_tma_kernel<false>, (size_t)TmaGramCfg::SMEM)) != cudaSuccess) return e;
    return kf_ensure_smem(ctx, kf_gram_tma_kernel<true>, (size_t)TmaGramCfg::SMEM);
}

}  // namespace

int kf_launch_gemm_tasks(kf_ctx* ctx, const KfGemmTask* d_tasks, int ntasks, bool weighted, cudaStream_t st) {
    if (ntasks <= 0) return KF_OK;
    KF_CUDA(ctx, ensure_attrs(ctx));
    if (weighted)
        kf_gram_tile_kernel<true><<<ntasks, KF_GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(d_tasks);
    else
        kf_gram_tile_kernel<false><<<ntasks, KF_GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(d_tasks);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}

int kf_launch_gram_tma(kf_ctx* ctx, const KfTmaTask* d_tasks, int ntasks, bool weighted, const CUtensorMap& tmap, cudaStream_t st) {
    if (ntasks <= 0) return KF_OK;
    KF_CUDA(ctx, ensure_attrs(ctx));
    if (weighted)
        kf_gram_tma_kernel<true><<<ntasks, TmaGramCfg::THREADS, TmaGramCfg::SMEM, st>>>(d_tasks, tmap);
    else
        kf_gram_tma_kernel<false><<<ntasks, TmaGramCfg::THREADS, TmaGramCfg::SMEM, st>>>(d_tasks, tmap);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}

int kf_launch_gemm_grid(kf_ctx* ctx, const KfGemmGrid& g, cudaStream_t st) {
    if (g.m <= 0 || g.n <= 0 || g.k1 <= g.k0) return KF_OK;
    KF_CUDA(ctx, ensure_attrs(ctx));
    dim3 grid((g.n + KF_CTA_N - 1) / KF_CTA_N, (g.m + KF_CTA_M - 1) / KF_CTA_M);
    if (g.ksplit > 1) {
        if (g.ksplit != kf_gemm_ksplit(g.k1 - g.k0, g.ksplit) || g.accumulate) {     // an empty chunk would leave its slab unwritten
            ctx->err = "kf_launch_gemm_grid: ksplit must come from kf_gemm_ksplit and accumulate must be 0";
            return KF_EINVAL;
        }
        grid.z = g.ksplit;
    }
    kf_gemm_grid_kernel<<<grid, KF_GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(g);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    double tiles = (double)grid.x * grid.y;
    if (g.lower_only) tiles *= 0.5;
    ctx->dmma_flops += tiles * 2.0 * KF_CTA_M * KF_CTA_N * (double)(g.k1 - g.k0);
    return KF_OK;
}

// out[m][n] (+)= alpha * sum_{k in [k0,k1)} A[m * lda + k] * B[k * ldb + n]   (B k-major; n a multiple of 64, k0/k1 of 16)
int kf_launch_gemm_bkmajor(kf_ctx* ctx, const KfGemmGrid& g, cudaStream_t st) {
    if (g.m <= 0 || g.n <= 0 || g.k1 <= g.k0) return KF_OK;
    if (g.n % KF_CTA_N || (g.k1 - g.k0) % KF_BK || (g.ldb & 1) || (reinterpret_cast<unsigned long long>(g.B) & 15ull)) {
        ctx->err = "kf_launch_gemm_bkmajor: n must be a multiple of 64, the k range of 16, B 16-byte aligned with an even ld";
        return KF_EINVAL;
    }
    KF_CUDA(ctx, ensure_attrs(ctx));
    dim3 grid(g.n / KF_CTA_N, (g.m + KF_CTA_M - 1) / KF_CTA_M);
    kf_gemm_bkmajor_kernel<<<grid, KF_GEMM_THREADS, kfg::bkmajor_smem_bytes<GramCfg>(), st>>>(g);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    ctx->dmma_flops += (double)grid.x * grid.y * 2.0 * KF_CTA_M * KF_CTA_N * (double)(g.k1 - g.k0);
    return KF_OK;
}
