// Pre-processing on the device (SURVEY §8f "next #3"): the raw merged training series crosses the boundary once and
// get_scale (Ksysid.m:180-229), get_zeta (868-907) and get_snapshotPairs (910-984) run as kernels that write the snapshot
// pairs alpha / beta / u (column-major, leading dimension M) straight into the buffers the lift reads.
//   scale:  offset = (max + min) / 2, factor = (max - min) / 2 (1 for a constant column), x -> (x - offset) / factor
//   zeta_i = [y_i, y_{i-1} .. y_{i-nd}, u_{i-1} .. u_{i-nd}], uzeta_i = u_i   (rows i >= nd of the MERGED series, as the
//            reference builds them: a delay window may straddle a trial boundary)
//   pairs:  candidate j pairs zeta_j with zeta_{j+1}; kept iff t_{nd+j} < t_{nd+j+1} (948); only the first
//           (kept - 1) are used (960); with snapshots = Inf all of those are taken (order is irrelevant to G, C, K).
#include <algorithm>

#include "kf_internal.h"

namespace {

constexpr int PP_THREADS = 256;
constexpr int PP_BLOCK = 2048;      // candidate pairs per CTA of the flag / scatter kernels

// grid (columns, parts): partial min / max of one column of [y | u]
__global__ void __launch_bounds__(PP_THREADS) kf_pp_minmax_kernel(const double* y, const double* u, long long T, int n, int m, int parts,
                                                                  double* pmin, double* pmax) {
    __shared__ double smin[PP_THREADS / 32], smax[PP_THREADS / 32];
    const int col = blockIdx.x, part = blockIdx.y;
    const double* x = col < n ? y + (long long)col * T : u + (long long)(col - n) * T;
    const long long per = (T + parts - 1) / parts, r0 = part * per, r1 = min(T, r0 + per);
    double mn = INFINITY, mx = -INFINITY;
    for (long long r = r0 + threadIdx.x; r < r1; r += PP_THREADS) {
        const double v = x[r];
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    for (int off = 16; off > 0; off >>= 1) {
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, off));
        mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < PP_THREADS / 32; ++w) { mn = fmin(mn, smin[w]); mx = fmax(mx, smax[w]); }
        pmin[col * parts + part] = mn;
        pmax[col * parts + part] = mx;
    }
}

// one thread per column: scale[col] = offset, scale[ncol + col] = factor   (Ksysid.m:193-204)
__global__ void kf_pp_scale_kernel(const double* pmin, const double* pmax, int ncol, int parts, double* scale) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    double mn = INFINITY, mx = -INFINITY;
    for (int p = 0; p < parts; ++p) { mn = fmin(mn, pmin[col * parts + p]); mx = fmax(mx, pmax[col * parts + p]); }
    const double half = (mx - mn) / 2.0;
    scale[col] = (mx + mn) / 2.0;
    scale[ncol + col] = half == 0.0 ? 1.0 : half;
}

// per CTA of PP_BLOCK candidates: number of kept pairs
__global__ void __launch_bounds__(PP_THREADS) kf_pp_count_kernel(const double* t, long long ncand, int nd, int* block_count) {
    __shared__ int wsum[PP_THREADS / 32];
    const long long j0 = (long long)blockIdx.x * PP_BLOCK;
    int c = 0;
    for (int e = threadIdx.x; e < PP_BLOCK; e += PP_THREADS) {
        const long long j = j0 + e;
        if (j < ncand && t[nd + j] < t[nd + j + 1]) ++c;
    }
    for (int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < PP_THREADS / 32; ++w) c += wsum[w];
        block_count[blockIdx.x] = c;
    }
}

// single CTA: exclusive prefix sum of the block counts (long long), total in block_off[nblocks]
__global__ void __launch_bounds__(1024) kf_pp_scan_kernel(const int* block_count, long long nblocks, long long* block_off) {
    __shared__ long long sh[1024];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long b0 = 0; b0 < nblocks; b0 += 1024) {
        const long long b = b0 + threadIdx.x;
        const long long v = b < nblocks ? block_count[b] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const long long add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        if (b < nblocks) block_off[b] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_off[nblocks] = carry;
}

struct PpArgs {
    const double* t; const double* y; const double* u; const double* scale;   // scale: [offset(n+m) | factor(n+m)]
    long long T, ncand, M;
    int n, m, nd;
    const long long* block_off;
    double* alpha; double* beta; double* uo;
};

// kept candidate with rank k < M writes row k of alpha, beta, u (scaled)
__global__ void __launch_bounds__(PP_THREADS) kf_pp_pairs_kernel(const PpArgs a) {
    __shared__ int wsum[PP_THREADS / 32];
    __shared__ long long base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long j0 = (long long)blockIdx.x * PP_BLOCK;
    if (tid == 0) base = a.block_off[blockIdx.x];
    __syncthreads();
    const int ncol = a.n + a.m;
    const double* off = a.scale;
    const double* fac = a.scale + ncol;
    for (int e0 = 0; e0 < PP_BLOCK; e0 += PP_THREADS) {
        const long long j = j0 + e0 + tid;
        const bool keep = j < a.ncand && a.t[a.nd + j] < a.t[a.nd + j + 1];
        const unsigned msk = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[warp] = __popc(msk);
        __syncthreads();
        long long k = base;
        for (int w = 0; w < warp; ++w) k += wsum[w];
        k += __popc(msk & ((1u << lane) - 1u));
        if (keep && k < a.M) {
            const long long i = a.nd + j;          // row of the merged series holding y_i of zeta_j
            // zeta = [y_i | y_{i-1} .. y_{i-nd} | u_{i-1} .. u_{i-nd}]
            int c = 0;
            for (int d = 0; d <= a.nd; ++d)
                for (int q = 0; q < a.n; ++q, ++c) {
                    const double* yq = a.y + (long long)q * a.T;
                    a.alpha[k + (long long)c * a.M] = (yq[i - d] - off[q]) / fac[q];
                    a.beta[k + (long long)c * a.M] = (yq[i + 1 - d] - off[q]) / fac[q];
                }
            for (int d = 1; d <= a.nd; ++d)
                for (int q = 0; q < a.m; ++q, ++c) {
                    const double* uq = a.u + (long long)q * a.T;
                    a.alpha[k + (long long)c * a.M] = (uq[i - d] - off[a.n + q]) / fac[a.n + q];
                    a.beta[k + (long long)c * a.M] = (uq[i + 1 - d] - off[a.n + q]) / fac[a.n + q];
                }
            for (int q = 0; q < a.m; ++q) a.uo[k + (long long)q * a.M] = (a.u[i + (long long)q * a.T] - off[a.n + q]) / fac[a.n + q];
        }
        __syncthreads();
        if (tid == 0) {
            int s = 0;
            for (int w = 0; w < PP_THREADS / 32; ++w) s += wsum[w];
            base += s;
        }
        __syncthreads();
    }
}

}  // namespace

// scale factors of the merged series (device): d_scale = [offset(n+m) | factor(n+m)]
int kf_pp_scale(kf_ctx* ctx, const double* d_y, const double* d_u, long long T, int n, int m, double* d_scale, cudaStream_t st) {
    const int ncol = n + m;
    const int parts = (int)std::max<long long>(1, std::min<long long>(4LL * ctx->sm_count / std::max(ncol, 1) + 1, (T + 4095) / 4096));
    KF_CUDA(ctx, ctx->d_tmp.ensure((size_t)2 * ncol * parts * sizeof(double)));
    double* pmin = ctx->d_tmp.as<double>();
    double* pmax = pmin + (size_t)ncol * parts;
    kf_pp_minmax_kernel<<<dim3(ncol, parts), PP_THREADS, 0, st>>>(d_y, d_u, T, n, m, parts, pmin, pmax);
    kf_pp_scale_kernel<<<(ncol + 63) / 64, 64, 0, st>>>(pmin, pmax, ncol, parts, d_scale);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return KF_OK;
}

// number of snapshot pairs the series yields: (kept candidates) - 1   (Ksysid.m:948, 960); leaves the block offsets in d_misc2
int kf_pp_count_pairs(kf_ctx* ctx, const double* d_t, long long T, int nd, long long* M_out, cudaStream_t st) {
    const long long ncand = T - nd - 1;
    if (ncand < 1) { *M_out = 0; return KF_OK; }
    const long long nblocks = (ncand + PP_BLOCK - 1) / PP_BLOCK;
    KF_CUDA(ctx, ctx->d_K3.ensure((size_t)nblocks * sizeof(int) + (size_t)(nblocks + 2) * sizeof(long long) + 16));
    long long* d_off = ctx->d_K3.as<long long>();
    int* d_cnt = reinterpret_cast<int*>(d_off + nblocks + 2);
    kf_pp_count_kernel<<<(unsigned)nblocks, PP_THREADS, 0, st>>>(d_t, ncand, nd, d_cnt);
    kf_pp_scan_kernel<<<1, 1024, 0, st>>>(d_cnt, nblocks, d_off);
    KF_CUDA(ctx, cudaGetLastError());
    long long kept = 0;
    KF_CUDA(ctx, cudaMemcpyAsync(&kept, d_off + nblocks, sizeof(long long), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 2;
    *M_out = std::max<long long>(kept - 1, 0);
    return KF_OK;
}

// scatter the M pairs (after kf_pp_scale and kf_pp_count_pairs on the same series)
int kf_pp_pairs(kf_ctx* ctx, const double* d_t, const double* d_y, const double* d_u, const double* d_scale, long long T, int n, int m,
                int nd, long long M, double* d_alpha, double* d_beta, double* d_uo, cudaStream_t st) {
    const long long ncand = T - nd - 1;
    const long long nblocks = (ncand + PP_BLOCK - 1) / PP_BLOCK;
    PpArgs a{};
    a.t = d_t; a.y = d_y; a.u = d_u; a.scale = d_scale;
    a.T = T; a.ncand = ncand; a.M = M; a.n = n; a.m = m; a.nd = nd;
    a.block_off = ctx->d_K3.as<long long>();
    a.alpha = d_alpha; a.beta = d_beta; a.uo = d_uo;
    kf_pp_pairs_kernel<<<(unsigned)nblocks, PP_THREADS, 0, st>>>(a);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}
