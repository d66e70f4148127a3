// L1-ball constrained least squares on the GPU — replaces Ksysid.solve_KoopmanQP
// (Ksysid.m:1095-1176: quadprog on the split variables x = [K+; K-] >= 0, 1'x <= t):
//
//     min_K  0.5 tr(K' G K) - tr(C' K)    s.t.  ||vec K||_1 <= t        (G = Px'Px, C = Px'Py)
//
// The columns of K couple only through the single budget t, so by the KKT conditions
// there is one multiplier lam >= 0 such that every column solves the penalised problem
// min 0.5 k'Gk - c_j'k + lam ||k||_1, and either (lam = 0, ||K||_1 <= t) or ||K||_1 = t.
//   * inner solve: cyclic coordinate descent in covariance form, ONE CTA PER COLUMN, the
//     column k_j and its gradient residual q_j = c_j - G k_j live in shared memory; all P
//     columns (and so all SMs) run concurrently with no inter-CTA synchronisation; glmnet-style
//     active-set sweeps;
//   * outer solve: bracketing + Illinois secant on phi(lam) = ||K(lam)||_1 - t (host-driven,
//     one tiny reduction per evaluation), warm-started;
//   * exact last step: on the final sign pattern K(lam) is affine in lam; D = dK/dlam is
//     obtained with the same kernel (restricted Gauss-Seidel on G_SS d = sign_S) and lam is
//     corrected so that ||K||_1 = t to rounding.
// A whole vector of budgets reuses G, C and warm-starts from the previous budget.
#include <algorithm>
#include <cmath>

#include "kf_internal.h"

namespace {

constexpr int CD_THREADS = 256;

struct CdArgs {
    const double* G; long long ldg;
    const double* dG;            // diag(G)
    const double* R; long long ldr;   // right-hand sides: C (mode 0/2) ; ignored in mode 1
    double* K; long long ldk;    // mode 0: warm start in / solution out; mode 1: D out; mode 2: read only
    const double* K0; long long ldk0;  // mode 1: support + signs
    int P;
    int col0;                    // first column handled (columns col0 .. col0+gridDim.x-1)
    int skip0, skip1;            // pinned columns [skip0, skip1) are left untouched
    double lam, tol;
    int max_sweeps, mode;        // 0 lasso CD, 1 restricted Gauss-Seidel for dK/dlam, 2 evaluate
    double* col_l1; double* col_obj; double* col_aux; int* col_iters;
};

__device__ __forceinline__ double soft(double x, double lam) {
    return x > lam ? x - lam : (x < -lam ? x + lam : 0.0);
}

__device__ __forceinline__ double cta_sum(double v, double* sh) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) sh[0] = v;
    }
    __syncthreads();
    return sh[0];
}

__global__ void __launch_bounds__(CD_THREADS) kf_cd_kernel(const CdArgs a) {
    extern __shared__ __align__(16) double cd_smem[];
    __shared__ double red[32];
    const int P = a.P, tid = threadIdx.x;
    const int col = a.col0 + blockIdx.x;
    if (col >= a.skip0 && col < a.skip1) return;
    double* k = cd_smem;            // current column
    double* q = cd_smem + P;        // q = rhs - G k
    int* act = reinterpret_cast<int*>(cd_smem + 2 * P);   // active index list
    __shared__ int s_nact;

    // ---- initialise k, rhs
    const double* Kcol = a.K + (long long)col * a.ldk;
    const double* K0col = a.mode == 1 ? a.K0 + (long long)col * a.ldk0 : nullptr;
    for (int i = tid; i < P; i += CD_THREADS) {
        if (a.mode == 1) {
            const double s0 = K0col[i];
            k[i] = 0.0;
            q[i] = s0 > 0.0 ? 1.0 : (s0 < 0.0 ? -1.0 : 0.0);   // rhs = sign pattern on the support
        } else {
            k[i] = Kcol[i];
            q[i] = a.R[(long long)col * a.ldr + i];
        }
    }
    __syncthreads();
    if (a.mode != 1) {   // q = c - G k for the warm start
        for (int i = 0; i < P; ++i) {
            const double ki = k[i];
            if (ki != 0.0) {
                const double* g = a.G + (long long)i * a.ldg;
                for (int r = tid; r < P; r += CD_THREADS) q[r] = fma(-ki, g[r], q[r]);
            }
        }
        __syncthreads();
    }

    int sweeps = 0;
    if (a.mode != 2) {
        // mode 1: the support is fixed to supp(K0)
        auto in_support = [&](int i) -> bool { return a.mode == 0 || K0col[i] != 0.0; };
        bool full = true;
        for (; sweeps < a.max_sweeps; ++sweeps) {
            double maxd = 0.0, maxk = 0.0;
            const int n = full ? P : s_nact;
            for (int ii = 0; ii < n; ++ii) {
                const int i = full ? ii : act[ii];
                const double gii = a.dG[i];
                if (!(gii > 0.0) || !in_support(i)) continue;
                const double ki = k[i];
                const double rho = fma(gii, ki, q[i]);
                const double nw = (a.mode == 0 ? soft(rho, a.lam) : rho) / gii;
                const double d = nw - ki;
                if (d != 0.0) {   // uniform across the CTA: every thread sees the same k, q
                    __syncthreads();          // all reads of q[i], k[i] done before they change
                    const double* g = a.G + (long long)i * a.ldg;
                    for (int r = tid; r < P; r += CD_THREADS) q[r] = fma(-d, g[r], q[r]);
                    if (tid == 0) k[i] = nw;
                    __syncthreads();
                    maxd = fmax(maxd, fabs(d));
                }
                maxk = fmax(maxk, fabs(nw));
            }
            const bool conv = maxd <= a.tol * fmax(maxk, 1e-300);
            if (full) {
                if (conv) { ++sweeps; break; }
                // rebuild the active list and iterate on it until it converges
                __syncthreads();
                if (tid == 0) {
                    int c = 0;
                    for (int i = 0; i < P; ++i)
                        if (k[i] != 0.0) act[c++] = i;
                    s_nact = c;
                }
                __syncthreads();
                full = false;
            } else if (conv) {
                full = true;   // verify with a full sweep (KKT on the inactive set)
            }
        }
    }

    // ---- outputs: column, ||k||_1, objective -0.5 k'(c + q)  (k'Gk = k'(c - q)), s'd for mode 1
    double l1 = 0.0, ob = 0.0, aux = 0.0;
    for (int i = tid; i < P; i += CD_THREADS) {
        const double ki = k[i];
        if (a.mode != 2) a.K[(long long)col * a.ldk + i] = ki;
        l1 += fabs(ki);
        if (a.mode == 1) {
            const double s0 = K0col[i];
            aux += (s0 > 0.0 ? ki : (s0 < 0.0 ? -ki : 0.0));
        } else {
            const double c = a.R[(long long)col * a.ldr + i];
            ob += -0.5 * ki * (c + q[i]);
        }
    }
    l1 = cta_sum(l1, red);
    ob = cta_sum(ob, red);
    aux = cta_sum(aux, red);
    if (tid == 0) {
        a.col_l1[col] = l1;
        a.col_obj[col] = ob;
        a.col_aux[col] = aux;
        a.col_iters[col] = sweeps;
    }
}

// single CTA: sums of the per-column outputs (fixed order -> deterministic); out[0..2] = l1, obj, aux, out[3] = max sweeps
__global__ void __launch_bounds__(1024) kf_cd_reduce_kernel(const double* l1, const double* ob, const double* aux,
                                                            const int* iters, int n, int skip0, int skip1, double* out) {
    __shared__ double red[32];
    double a = 0, b = 0, c = 0, m = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (i >= skip0 && i < skip1) continue;
        a += l1[i]; b += ob[i]; c += aux[i]; m = fmax(m, (double)iters[i]);
    }
    a = cta_sum(a, red);
    b = cta_sum(b, red);
    c = cta_sum(c, red);
    // max via sum trick is wrong; do a proper max reduction
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double mm = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) mm = fmax(mm, red[w]);
        out[0] = a; out[1] = b; out[2] = c; out[3] = mm;
    }
}

__global__ void kf_diag_kernel(const double* G, long long ld, int P, double shift, double* G2, double* dG) {
    // optional G2 = G + shift I (in place allowed), dG = diag
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const double d = G[(long long)i * ld + i] + shift;
    if (G2) G2[(long long)i * ld + i] = d;
    dG[i] = d;
}

// K = K0 - dl * D on columns outside [skip0, skip1); flags sign flips
__global__ void kf_axpy_sign_kernel(double* K, const double* D, long long ld, int P, int ncols, int skip0, int skip1, double dl,
                                    int* flips) {
    const long long n = (long long)P * ncols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % P), c = (int)(e / P);
        if (c >= skip0 && c < skip1) continue;
        const double k0 = K[(long long)c * ld + i];
        if (k0 == 0.0) continue;
        const double k1 = k0 - dl * D[(long long)c * ld + i];
        if ((k1 > 0.0) != (k0 > 0.0)) atomicAdd(flips, 1);
        K[(long long)c * ld + i] = k1;
    }
}

__global__ void kf_scale_kernel(double* K, long long ld, int P, int ncols, int skip0, int skip1, double s) {
    const long long n = (long long)P * ncols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % P), c = (int)(e / P);
        if (c >= skip0 && c < skip1) continue;
        K[(long long)c * ld + i] *= s;
    }
}

__global__ void kf_absmax_kernel(const double* C, long long ld, int P, int ncols, int skip0, int skip1, double* out) {
    // single CTA
    __shared__ double red[32];
    double m = 0;
    const long long n = (long long)P * ncols;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
        const int i = (int)(e % P), c = (int)(e / P);
        if (c >= skip0 && c < skip1) continue;
        m = fmax(m, fabs(C[(long long)c * ld + i]));
    }
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double mm = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) mm = fmax(mm, red[w]);
        out[0] = mm;
    }
}

// the dynamic shared-memory limit of the CD kernel only ever grows (a smaller value would make a later,
// larger launch fail with "invalid argument")
int ensure_cd_smem(kf_ctx* ctx, size_t smem) {
    static size_t smem_set = 48 * 1024;
    if (smem > smem_set) {
        KF_CUDA(ctx, cudaFuncSetAttribute(kf_cd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    return KF_OK;
}

}  // namespace

// Solve one ACTIVE budget t (the caller has checked that the unconstrained minimiser violates it).
// G (possibly shifted), C: Pp-strided P x P.  K: warm start in / solution out.  Columns [fix_c0, fix_c1) hold the
// pinned delay pattern (already in K); `t` is the budget left for the free columns.
// lam_start > 0: an upper bracket from the previous (smaller) budget of an ascending sweep, with K = K(lam_start)
// and phi_start = ||K||_1 - t < 0; lam_start <= 0: cold start from lam_max with K = 0.
int kf_solve_l1ball(kf_ctx* ctx, int P, int Pp, const double* G, const double* C, double t, int fix_c0, int fix_c1,
                    double lam_start, double phi_start, int max_iter, double tol, double* K, KfQpResult* res, cudaStream_t st) {
    const long long ld = Pp;
    const size_t smem = (size_t)P * (2 * sizeof(double) + sizeof(int)) + 16;
    KF_TRY(ensure_cd_smem(ctx, smem));
    // scratch: dG[P] | l1[P] | obj[P] | aux[P] | out[4] | iters[P] | flips
    KF_CUDA(ctx, ctx->d_K3.ensure((size_t)(4 * Pp + 8) * sizeof(double) + (size_t)(Pp + 4) * sizeof(int)));
    double* dG = ctx->d_K3.as<double>();
    double* c_l1 = dG + Pp;
    double* c_ob = c_l1 + Pp;
    double* c_aux = c_ob + Pp;
    double* d_out = c_aux + Pp;
    int* c_it = reinterpret_cast<int*>(d_out + 8);
    int* d_flips = c_it + Pp;
    KF_CUDA(ctx, ctx->d_K2.ensure((size_t)Pp * Pp * sizeof(double)));   // D = dK/dlam
    double* D = ctx->d_K2.as<double>();

    kf_diag_kernel<<<(P + 255) / 256, 256, 0, st>>>(G, ld, P, 0.0, nullptr, dG);
    ctx->launches += 1;

    // inner sweeps per evaluation are bounded so that an ill-conditioned Gram cannot run away; a column that
    // hits the bound is reported through res->capped
    const int max_sweeps = max_iter > 0 ? max_iter : (P <= 256 ? 100000 : 2000);
    const double cd_tol = tol > 0 ? tol : 1e-13;
    int evals = 0, capped = 0;
    double h[4];
    auto run_cd = [&](int mode, double lam, double* Kout, const double* K0) -> int {
        CdArgs a{};
        a.G = G; a.ldg = ld; a.dG = dG; a.R = C; a.ldr = ld;
        a.K = Kout; a.ldk = ld; a.K0 = K0; a.ldk0 = ld;
        a.P = P; a.col0 = 0; a.skip0 = fix_c0; a.skip1 = fix_c1;
        a.lam = lam; a.tol = cd_tol; a.max_sweeps = max_sweeps; a.mode = mode;
        a.col_l1 = c_l1; a.col_obj = c_ob; a.col_aux = c_aux; a.col_iters = c_it;
        kf_cd_kernel<<<P, CD_THREADS, smem, st>>>(a);
        kf_cd_reduce_kernel<<<1, 1024, 0, st>>>(c_l1, c_ob, c_aux, c_it, P, fix_c0, fix_c1, d_out);
        KF_CUDA(ctx, cudaGetLastError());
        KF_CUDA(ctx, cudaMemcpyAsync(h, d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->launches += 2;
        ++evals;
        if (mode != 2 && h[3] >= max_sweeps) ++capped;
        return KF_OK;
    };

    double lam_hi, phi_hi, lam_lo = 0, phi_lo = 0, lam;
    if (lam_start > 0) {
        lam_hi = lam_start;
        phi_hi = phi_start;
    } else {
        kf_absmax_kernel<<<1, 1024, 0, st>>>(C, ld, P, P, fix_c0, fix_c1, d_out);
        double lam_max = 0;
        KF_CUDA(ctx, cudaMemcpyAsync(&lam_max, d_out, sizeof(double), cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->launches += 1;
        lam_hi = lam_max;   // K(lam_max) = 0
        phi_hi = -t;
        if (fix_c0 > 0) KF_CUDA(ctx, cudaMemset2DAsync(K, ld * sizeof(double), 0, (size_t)P * sizeof(double), fix_c0, st));
        if (fix_c1 < P)   // the pinned columns [fix_c0, fix_c1) keep their pattern
            KF_CUDA(ctx, cudaMemset2DAsync(K + (size_t)std::max(fix_c1, 0) * ld, ld * sizeof(double), 0, (size_t)P * sizeof(double),
                                           P - std::max(fix_c1, 0), st));
    }
    // 1. bracket: shrink lam geometrically (warm-started) until ||K||_1 > t
    lam = lam_hi;
    bool bracket = false;
    for (int it = 0; it < 200; ++it) {
        lam *= 0.5;
        KF_TRY(run_cd(0, lam, K, nullptr));
        const double phi = h[0] - t;
        if (phi > 0) { lam_lo = lam; phi_lo = phi; bracket = true; break; }
        lam_hi = lam; phi_hi = phi;
        if (lam < 1e-300) break;
    }
    if (bracket) {
        // 2. Illinois secant on phi(lam) = ||K(lam)||_1 - t; the exact step below removes the remaining error,
        //    so a modest tolerance is enough here
        int side = 0;
        for (int it = 0; it < 60; ++it) {
            lam = (lam_lo * phi_hi - lam_hi * phi_lo) / (phi_hi - phi_lo);
            if (!(lam > lam_lo && lam < lam_hi)) lam = 0.5 * (lam_lo + lam_hi);
            KF_TRY(run_cd(0, lam, K, nullptr));
            const double phi = h[0] - t;
            if (fabs(phi) <= 1e-10 * t) break;
            if (phi > 0) {
                lam_lo = lam; phi_lo = phi;
                if (side == 1) phi_hi *= 0.5;
                side = 1;
            } else {
                lam_hi = lam; phi_hi = phi;
                if (side == -1) phi_lo *= 0.5;
                side = -1;
            }
            if (lam_hi - lam_lo <= 1e-14 * lam_hi) break;
        }
        // 3. exact last step on the fixed sign pattern: D = dK/dlam, ||K(lam + dl)||_1 = ||K||_1 - dl * sum s'd
        const double l1_0 = h[0];
        KF_TRY(run_cd(1, 0.0, D, K));
        const double den = h[2];
        if (den > 0) {
            const double dl = (l1_0 - t) / den;
            KF_CUDA(ctx, cudaMemsetAsync(d_flips, 0, sizeof(int), st));
            KF_CUDA(ctx, ctx->d_tmp.ensure((size_t)Pp * Pp * sizeof(double)));
            double* Kbak = ctx->d_tmp.as<double>();
            KF_CUDA(ctx, cudaMemcpyAsync(Kbak, K, (size_t)Pp * P * sizeof(double), cudaMemcpyDeviceToDevice, st));
            kf_axpy_sign_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(K, D, ld, P, P, fix_c0, fix_c1, dl, d_flips);
            int flips = 0;
            KF_CUDA(ctx, cudaMemcpyAsync(&flips, d_flips, sizeof(int), cudaMemcpyDeviceToHost, st));
            KF_CUDA(ctx, cudaStreamSynchronize(st));
            if (flips) KF_CUDA(ctx, cudaMemcpyAsync(K, Kbak, (size_t)Pp * P * sizeof(double), cudaMemcpyDeviceToDevice, st));
            else lam += dl;
        }
    }
    // 4. evaluate; enforce feasibility to rounding
    KF_TRY(run_cd(2, 0.0, K, nullptr));
    if (h[0] > t && h[0] > 0) {
        kf_scale_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(K, ld, P, P, fix_c0, fix_c1, t / h[0]);
        KF_TRY(run_cd(2, 0.0, K, nullptr));
    }
    res->objective = h[1];
    res->l1 = h[0];
    res->iters = evals;
    res->lam = lam;
    res->capped = capped;
    return KF_OK;
}

// G += shift * I  (the PSD conditioning branch, Ksysid.m:1119)
int kf_add_diag(kf_ctx* ctx, double* G, int Pp, int P, double shift, cudaStream_t st) {
    KF_CUDA(ctx, ctx->d_K3.ensure((size_t)(4 * Pp + 8) * sizeof(double) + (size_t)(Pp + 4) * sizeof(int)));
    kf_diag_kernel<<<(P + 255) / 256, 256, 0, st>>>(G, Pp, P, shift, G, ctx->d_K3.as<double>());
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}

// objective 0.5 tr(K'GK) - tr(C'K) and ||vec K||_1 over all P columns
int kf_qp_evaluate(kf_ctx* ctx, int P, int Pp, const double* G, const double* C, const double* K, KfQpResult* res, cudaStream_t st) {
    const long long ld = Pp;
    const size_t smem = (size_t)P * (2 * sizeof(double) + sizeof(int)) + 16;
    KF_TRY(ensure_cd_smem(ctx, smem));
    KF_CUDA(ctx, ctx->d_K3.ensure((size_t)(4 * Pp + 8) * sizeof(double) + (size_t)(Pp + 4) * sizeof(int)));
    double* dG = ctx->d_K3.as<double>();
    double* c_l1 = dG + Pp;
    double* c_ob = c_l1 + Pp;
    double* c_aux = c_ob + Pp;
    double* d_out = c_aux + Pp;
    int* c_it = reinterpret_cast<int*>(d_out + 8);
    kf_diag_kernel<<<(P + 255) / 256, 256, 0, st>>>(G, ld, P, 0.0, nullptr, dG);
    CdArgs a{};
    a.G = G; a.ldg = ld; a.dG = dG; a.R = C; a.ldr = ld;
    a.K = const_cast<double*>(K); a.ldk = ld; a.K0 = nullptr; a.ldk0 = ld;
    a.P = P; a.col0 = 0; a.skip0 = 0; a.skip1 = 0;
    a.lam = 0; a.tol = 0; a.max_sweeps = 0; a.mode = 2;
    a.col_l1 = c_l1; a.col_obj = c_ob; a.col_aux = c_aux; a.col_iters = c_it;
    kf_cd_kernel<<<P, CD_THREADS, smem, st>>>(a);
    kf_cd_reduce_kernel<<<1, 1024, 0, st>>>(c_l1, c_ob, c_aux, c_it, P, 0, 0, d_out);
    KF_CUDA(ctx, cudaGetLastError());
    double h[4];
    KF_CUDA(ctx, cudaMemcpyAsync(h, d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 3;
    res->l1 = h[0];
    res->objective = h[1];
    return KF_OK;
}
