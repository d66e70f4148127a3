// Principal components of the lifted snapshots on the device (replaces MATLAB's pca at Ksysid.m:1498, the `dim_red` branch of
// get_econ_observables, Ksysid.m:1495-1517).
//
//   pass 1: lift the points chunk by chunk (lift.cu) and sum every feature            -> mu
//   pass 2: lift again, subtract mu IN the lifted chunk, Gc += Psi_c Psi_c' (DMMA GEMM, two-level summation over the points)
//           — the covariance is formed from CENTRED data like pca's SVD, not as G - M mu mu' (that cancellation cost the round-1
//           path 4-5 digits on features whose mean dominates their spread)
//   eig   : one-sided (Hestenes) Jacobi on the symmetric Gc: W = Gc, V = I; a round of the round-robin tournament rotates n/2
//           disjoint column pairs of W and V in one launch (a CTA per pair: three dot products, one plane rotation); converged
//           when a whole sweep rotates nothing; eigenvalue j = v_j' w_j, eigenvector j = v_j.
// Sorting by decreasing eigenvalue and MATLAB's sign convention (largest |component| of a vector positive) happen on the n values /
// n x n matrix after the copy-out.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "kf_internal.h"

namespace {

constexpr int PCA_THREADS = 128;

__device__ __forceinline__ double pca_block_sum(double v, double* sh) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) s += sh[q];
    return s;
}

// sums[row] += sum of the chunk's row (one CTA per feature; chunks arrive in stream order: deterministic)
__global__ void __launch_bounds__(256) kf_pca_rowsum_kernel(const double* __restrict__ Psi, long long ld, long long cnt, double* sums) {
    __shared__ double sh[8];
    const double* r = Psi + (long long)blockIdx.x * ld;
    double s = 0.0;
    for (long long i = threadIdx.x; i < cnt; i += blockDim.x) s += r[i];
    s = pca_block_sum(s, sh);
    if (threadIdx.x == 0) sums[blockIdx.x] += s;
}
// Psi(row, 0:cnt) -= mu(row); the k padding [cnt, kpad) is zeroed
__global__ void __launch_bounds__(256) kf_pca_center_kernel(double* Psi, long long ld, long long cnt, long long kpad, const double* __restrict__ sums,
                                                            double inv_rows) {
    double* r = Psi + (long long)blockIdx.x * ld;
    const double mu = sums[blockIdx.x] * inv_rows;
    for (long long i = threadIdx.x; i < kpad; i += blockDim.x) r[i] = i < cnt ? r[i] - mu : 0.0;
}
__global__ void kf_pca_axpy_kernel(const double* __restrict__ x, double* y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += x[i];
}
// W = scale * sym(Gc) from its lower triangle, V = I  (n2 x n2; the padding row / column of an odd n stays zero)
__global__ void kf_pca_init_kernel(const double* __restrict__ Gc, int n, int n2, double scale, double* W, double* V) {
    const long long tot = (long long)n2 * n2;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % n2), j = (int)(e / n2);
        double w = 0.0;
        if (i < n && j < n) w = scale * (i >= j ? Gc[(long long)j * n2 + i] : Gc[(long long)i * n2 + j]);
        W[e] = w;
        V[e] = (i == j) ? 1.0 : 0.0;
    }
}
// one round of the tournament: CTA b rotates the column pair (p, q) of W and V so that w_p' w_q = 0
__global__ void __launch_bounds__(PCA_THREADS) kf_pca_jacobi_kernel(double* W, double* V, int n2, int round, double tol, int* nrot) {
    __shared__ double sh[PCA_THREADS / 32];
    const int b = blockIdx.x, nm1 = n2 - 1;
    const int p = b == 0 ? nm1 : (round + b) % nm1;
    const int q = b == 0 ? round : (round - b + nm1) % nm1;
    double* wp = W + (long long)p * n2;
    double* wq = W + (long long)q * n2;
    double a = 0.0, bb = 0.0, g = 0.0;
    for (int i = threadIdx.x; i < n2; i += PCA_THREADS) {
        const double x = wp[i], y = wq[i];
        a = fma(x, x, a); bb = fma(y, y, bb); g = fma(x, y, g);
    }
    a = pca_block_sum(a, sh); bb = pca_block_sum(bb, sh); g = pca_block_sum(g, sh);
    if (!(fabs(g) > tol * sqrt(a * bb)) || a == 0.0 || bb == 0.0) return;     // uniform: every thread holds the same sums
    const double zeta = (bb - a) / (2.0 * g);
    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
    double* vp = V + (long long)p * n2;
    double* vq = V + (long long)q * n2;
    for (int i = threadIdx.x; i < n2; i += PCA_THREADS) {
        const double x = wp[i], y = wq[i];
        wp[i] = c * x - s * y; wq[i] = s * x + c * y;
        const double u = vp[i], v = vq[i];
        vp[i] = c * u - s * v; vq[i] = s * u + c * v;
    }
    if (threadIdx.x == 0) atomicAdd(nrot, 1);
}
// eigenvalue j = v_j' w_j (w_j = Gc v_j)
__global__ void __launch_bounds__(PCA_THREADS) kf_pca_rayleigh_kernel(const double* __restrict__ W, const double* __restrict__ V, int n2, double* lam) {
    __shared__ double sh[PCA_THREADS / 32];
    const double* w = W + (long long)blockIdx.x * n2;
    const double* v = V + (long long)blockIdx.x * n2;
    double s = 0.0;
    for (int i = threadIdx.x; i < n2; i += PCA_THREADS) s = fma(w[i], v[i], s);
    s = pca_block_sum(s, sh);
    if (threadIdx.x == 0) lam[blockIdx.x] = s;
}

}  // namespace

// d_V: rows x nv device points (ld = rows); the dictionary of ctx->prog (no dim_red).  mu, latent: n_full; coeff: n_full x n_full
// column-major (host buffers).  latent in decreasing order, coeff(:, j) the matching unit eigenvector.
int kf_pca_points(kf_ctx* ctx, long long rows, const double* d_V, double* mu, double* latent, double* coeff) {
    const KfProgram& p = ctx->prog;
    const int n = p.n_full();
    const int n2 = (n + 1) & ~1;                        // the tournament needs an even number of columns
    cudaStream_t st = ctx->stream;
    if (rows < 2 || p.n_pcs) {
        ctx->err = "kf_pca: at least two points and a dictionary without dim_red are required";
        return KF_EINVAL;
    }
    const long long chunk = std::min<long long>(kf_roundup(rows, KF_BK), 32768);
    const int kmax = 64;
    const size_t mat = (size_t)n2 * n2;
    // lifted chunk | sums | Gc | split-K slabs | W | V | lambda | rotation counter
    KF_CUDA(ctx, ctx->d_pca.ensure(((size_t)n * chunk + n2 + mat * (3 + kmax) + n2 + 16) * sizeof(double)));
    double* Psi = ctx->d_pca.as<double>();
    double* sums = Psi + (size_t)n * chunk;
    double* Gc = sums + n2;
    double* slabs = Gc + mat;
    double* W = slabs + mat * kmax;
    double* V = W + mat;
    double* lam = V + mat;
    int* nrot = reinterpret_cast<int*>(lam + n2);
    KF_CUDA(ctx, cudaMemsetAsync(sums, 0, (n2 + mat) * sizeof(double), st));          // sums and Gc
    const long long nch = (rows + chunk - 1) / chunk;
    for (int pass = 0; pass < 2; ++pass)
        for (long long c = 0; c < nch; ++c) {
            const long long c0 = c * chunk, cnt = std::min(chunk, rows - c0);
            KF_TRY(kf_launch_lift_points_chunk(ctx, ctx->d_ops.as<KfOp>(), ctx->d_centres.as<double>(), p.nv, n, d_V, rows, c0, cnt, Psi, chunk, st));
            if (pass == 0) {
                kf_pca_rowsum_kernel<<<n, 256, 0, st>>>(Psi, chunk, cnt, sums);
            } else {
                const long long kpad = kf_roundup(cnt, KF_BK);
                kf_pca_center_kernel<<<n, 256, 0, st>>>(Psi, chunk, cnt, kpad, sums, 1.0 / (double)rows);
                KfGemmGrid g{};
                g.A = Psi; g.lda = chunk; g.B = Psi; g.ldb = chunk; g.out = slabs; g.ldm = 1; g.ldn = n2;
                g.m = n; g.n = n; g.k0 = 0; g.k1 = (int)kpad; g.alpha = 1.0; g.accumulate = 0; g.lower_only = 1;
                g.ksplit = kf_gemm_ksplit((int)kpad, (int)std::min<long long>(kmax, std::max<long long>(1, kpad / 256)));
                g.slab = (long long)mat;
                if (g.ksplit > 1) {
                    KF_CUDA(ctx, cudaMemsetAsync(slabs, 0, mat * g.ksplit * sizeof(double), st));     // lower_only skips the upper tiles
                    KF_TRY(kf_launch_gemm_grid(ctx, g, st));
                    KF_TRY(kf_reduce_slabs(ctx, slabs, g.slab, g.ksplit, st));
                } else {
                    g.ksplit = 0;
                    KF_CUDA(ctx, cudaMemsetAsync(slabs, 0, mat * sizeof(double), st));
                    KF_TRY(kf_launch_gemm_grid(ctx, g, st));
                }
                kf_pca_axpy_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(slabs, Gc, (long long)mat);
            }
            KF_CUDA(ctx, cudaGetLastError());
            ctx->launches += 2;
        }
    // covariance = Gc / (rows - 1); Jacobi sweeps
    kf_pca_init_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(Gc, n, n2, 1.0 / (double)(rows - 1), W, V);
    int sweeps = 0;
    const double jtol = 4.0 * 2.220446049250313e-16 * std::sqrt((double)n2);     // a dot product of n terms is not resolved below this
    for (; sweeps < 40; ++sweeps) {
        KF_CUDA(ctx, cudaMemsetAsync(nrot, 0, sizeof(int), st));
        for (int r = 0; r < n2 - 1; ++r) kf_pca_jacobi_kernel<<<n2 / 2, PCA_THREADS, 0, st>>>(W, V, n2, r, jtol, nrot);
        KF_CUDA(ctx, cudaGetLastError());
        ctx->launches += n2 - 1;
        int h = 0;
        KF_CUDA(ctx, cudaMemcpyAsync(&h, nrot, sizeof(int), cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        if (h == 0) break;
    }
    kf_pca_rayleigh_kernel<<<n2, PCA_THREADS, 0, st>>>(W, V, n2, lam);
    KF_CUDA(ctx, cudaGetLastError());
    std::vector<double> hl(n2), hv(mat), hs(n2);
    KF_CUDA(ctx, cudaMemcpyAsync(hl.data(), lam, n2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaMemcpyAsync(hv.data(), V, mat * sizeof(double), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaMemcpyAsync(hs.data(), sums, n2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    // the padding column of an odd n is the unit vector e_n with eigenvalue 0: drop it; order by decreasing eigenvalue
    std::vector<int> ord;
    for (int j = 0; j < n2; ++j) {
        bool pad = false;
        if (n2 != n) pad = std::fabs(hv[(size_t)j * n2 + n]) > 0.5;
        if (!pad) ord.push_back(j);
    }
    if ((int)ord.size() != n) {
        ctx->err = "kf_pca: internal error (padding column not isolated)";
        return KF_ECUDA;
    }
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return hl[x] > hl[y]; });
    for (int j = 0; j < n; ++j) {
        const double* v = hv.data() + (size_t)ord[j] * n2;
        int imax = 0;
        for (int i = 1; i < n; ++i)
            if (std::fabs(v[i]) > std::fabs(v[imax])) imax = i;
        const double sg = v[imax] < 0.0 ? -1.0 : 1.0;      // MATLAB pca: the largest component of every vector is positive
        if (latent) latent[j] = std::max(hl[ord[j]], 0.0);
        if (coeff)
            for (int i = 0; i < n; ++i) coeff[(size_t)j * n + i] = sg * v[i];
    }
    if (mu)
        for (int i = 0; i < n; ++i) mu[i] = hs[i] / (double)rows;
    ctx->last_pca_sweeps = sweeps + 1;
    return KF_OK;
}
