// L1-ball constrained least squares on the GPU — replaces Ksysid.solve_KoopmanQP
// (Ksysid.m:1095-1176: quadprog on the split variables x = [K+; K-] >= 0, 1'x <= t):
//
//     min_K  0.5 tr(K' G K) - tr(C' K)    s.t.  ||vec K||_1 <= t        (G = Px'Px, C = Px'Py)
//
// The columns of K couple only through the single budget t, so by the KKT conditions
// there is one multiplier lam >= 0 such that every column solves the penalised problem
// min 0.5 k'Gk - c_j'k + lam ||k||_1, and either (lam = 0, ||K||_1 <= t) or ||K||_1 = t.
//   * inner solve: cyclic coordinate descent in covariance form, ONE CTA PER (COLUMN, BUDGET): the
//     column k_j and its gradient residual q_j = c_j - G k_j live in shared memory; all columns of
//     ALL budgets of a lasso vector run concurrently (grid = P x nt) with no inter-CTA
//     synchronisation; glmnet-style active-set sweeps;
//   * outer solve: every budget runs its own bracketing + Illinois secant on
//     phi_b(lam) = ||K_b(lam)||_1 - t_b, in lockstep (one launch evaluates all unfinished budgets,
//     one tiny per-budget reduction), warm-started from its previous iterate;
//   * exact last step: on the final sign pattern K(lam) is affine in lam; D = dK/dlam comes from
//     the same kernel (restricted Gauss-Seidel on G_SS d = sign_S) and lam is corrected so that
//     ||K||_1 = t to rounding.
#include <algorithm>
#include <cmath>

#include "kf_internal.h"

namespace {

constexpr int CD_THREADS = 256;

struct CdArgs {
    const double* G; long long ldg;
    const double* dG;                 // diag(G)
    const double* R; long long ldr;   // right-hand sides C (mode 0/2)
    double* K; long long ldk;         // budget b: K + b * kstride.  mode 0: warm start in / solution out; 1: D out; 2: read only
    const double* K0; long long ldk0; // mode 1: support + signs (budget b: K0 + b * kstride)
    long long kstride;
    int P;
    int skip0, skip1;                 // pinned columns [skip0, skip1) are left untouched
    const double* lam;                // per budget
    const int* active;                // per budget: 0 = skip this launch
    double tol;
    int max_sweeps, mode;             // 0 lasso CD, 1 restricted Gauss-Seidel for dK/dlam, 2 evaluate
    double* col_l1; double* col_obj; double* col_aux; int* col_iters;   // [budget][column]
    double* col_gmax;                 // mode 2: max_i |grad_i| of the column (duality-gap certificate)
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
    __shared__ int s_nact;
    const int P = a.P, tid = threadIdx.x;
    const int col = blockIdx.x, bud = blockIdx.y;
    if (!a.active[bud]) return;
    if (col >= a.skip0 && col < a.skip1) return;
    double* k = cd_smem;            // current column
    double* q = cd_smem + P;        // q = rhs - G k
    int* act = reinterpret_cast<int*>(cd_smem + 2 * P);   // active index list
    const double lam = a.lam[bud];

    // ---- initialise k, rhs
    double* Kcol = a.K + bud * a.kstride + (long long)col * a.ldk;
    const double* K0col = a.mode == 1 ? a.K0 + bud * a.kstride + (long long)col * a.ldk0 : nullptr;
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
                const double nw = (a.mode == 0 ? soft(rho, lam) : rho) / gii;
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
    double l1 = 0.0, ob = 0.0, aux = 0.0, gmax = 0.0;
    for (int i = tid; i < P; i += CD_THREADS) {
        const double ki = k[i];
        if (a.mode != 2) Kcol[i] = ki;
        l1 += fabs(ki);
        if (a.mode == 1) {
            const double s0 = K0col[i];
            aux += (s0 > 0.0 ? ki : (s0 < 0.0 ? -ki : 0.0));
        } else {
            const double c = a.R[(long long)col * a.ldr + i];
            ob += -0.5 * ki * (c + q[i]);
            if (a.mode == 2) {                // q = -grad: <grad, k> and ||grad||_inf for the Frank-Wolfe gap
                aux -= q[i] * ki;
                gmax = fmax(gmax, fabs(q[i]));
            }
        }
    }
    l1 = cta_sum(l1, red);
    ob = cta_sum(ob, red);
    aux = cta_sum(aux, red);
    if (a.mode == 2 && a.col_gmax) {
        for (int off = 16; off > 0; off >>= 1) gmax = fmax(gmax, __shfl_down_sync(0xffffffffu, gmax, off));
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = gmax;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < CD_THREADS / 32; ++w) gmax = fmax(gmax, red[w]);
            a.col_gmax[(long long)bud * P + col] = gmax;
        }
    }
    if (tid == 0) {
        const long long o = (long long)bud * P + col;
        a.col_l1[o] = l1;
        a.col_obj[o] = ob;
        a.col_aux[o] = aux;
        a.col_iters[o] = sweeps;
    }
}

// one CTA per budget: sums of the per-column outputs (fixed order -> deterministic);
// out[b][0..2] = l1, obj, aux, out[b][3] = max sweeps
__global__ void __launch_bounds__(256) kf_cd_reduce_kernel(const double* l1, const double* ob, const double* aux, const int* iters,
                                                           int n, int skip0, int skip1, const int* active, double* out) {
    __shared__ double red[32];
    const int bud = blockIdx.x;
    if (!active[bud]) return;
    const long long base = (long long)bud * n;
    double a = 0, b = 0, c = 0, m = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (i >= skip0 && i < skip1) continue;
        a += l1[base + i]; b += ob[base + i]; c += aux[base + i]; m = fmax(m, (double)iters[base + i]);
    }
    a = cta_sum(a, red);
    b = cta_sum(b, red);
    c = cta_sum(c, red);
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double mm = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) mm = fmax(mm, red[w]);
        out[bud * 4 + 0] = a; out[bud * 4 + 1] = b; out[bud * 4 + 2] = c; out[bud * 4 + 3] = mm;
    }
}

// single CTA: Frank-Wolfe duality gap pieces over the free columns: out[0] = sum <grad_j, k_j>, out[1] = max |grad|
__global__ void __launch_bounds__(256) kf_gap_reduce_kernel(const double* aux, const double* gmax, int n, int skip0, int skip1, int own_lo,
                                                            int own_hi, double* out) {
    __shared__ double red[32];
    double a = 0, m = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if ((i >= skip0 && i < skip1) || i < own_lo || i >= own_hi) continue;
        a += aux[i];
        m = fmax(m, gmax[i]);
    }
    a = cta_sum(a, red);
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double mm = 0;
        for (int w = 0; w < (blockDim.x >> 5); ++w) mm = fmax(mm, red[w]);
        out[0] = a; out[1] = mm;
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

// budget b = blockIdx.y.  apply == 0: count sign flips of K - dl_b * D;  apply == 1: K -= dl_b * D where allowed[b]
__global__ void kf_axpy_sign_kernel(double* K, const double* D, long long kstride, long long ld, int P, int ncols, int skip0, int skip1,
                                    const double* dl, const int* allowed, int* flips, int apply) {
    const int b = blockIdx.y;
    if (!allowed[b]) return;
    if (apply && flips[b]) return;
    double* Kb = K + b * kstride;
    const double* Db = D + b * kstride;
    const double d = dl[b];
    const long long n = (long long)P * ncols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % P), c = (int)(e / P);
        if (c >= skip0 && c < skip1) continue;
        const double k0 = Kb[(long long)c * ld + i];
        if (k0 == 0.0) continue;
        const double k1 = k0 - d * Db[(long long)c * ld + i];
        if (!apply) {
            if ((k1 > 0.0) != (k0 > 0.0)) atomicAdd(flips + b, 1);
        } else {
            Kb[(long long)c * ld + i] = k1;
        }
    }
}

__global__ void kf_scale_kernel(double* K, long long kstride, long long ld, int P, int ncols, int skip0, int skip1, const double* s) {
    const int b = blockIdx.y;
    const double sc = s[b];
    if (sc == 1.0) return;
    double* Kb = K + b * kstride;
    const long long n = (long long)P * ncols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % P), c = (int)(e / P);
        if (c >= skip0 && c < skip1) continue;
        Kb[(long long)c * ld + i] *= sc;
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
    KF_CUDA(ctx, kf_ensure_smem(ctx, kf_cd_kernel, smem));
    return KF_OK;
}

struct Scratch {
    double *dG, *c_l1, *c_ob, *c_aux, *c_gmax, *d_out, *d_lam, *d_dl;
    int *c_it, *d_active, *d_flips;
};

// scratch for nb budgets: dG[Pp] | l1,obj,aux,gmax [nb*Pp] | out[4 nb] | lam[nb] | dl[nb] | iters[nb*Pp] | active[nb] | flips[nb]
int carve(kf_ctx* ctx, int Pp, int nb, Scratch* s) {
    const size_t nd = (size_t)Pp + 4ull * nb * Pp + 6ull * nb + 8;
    const size_t ni = (size_t)nb * Pp + 2ull * nb + 8;
    KF_CUDA(ctx, ctx->d_K3.ensure(nd * sizeof(double) + ni * sizeof(int)));
    double* p = ctx->d_K3.as<double>();
    s->dG = p; p += Pp;
    s->c_l1 = p; p += (size_t)nb * Pp;
    s->c_ob = p; p += (size_t)nb * Pp;
    s->c_aux = p; p += (size_t)nb * Pp;
    s->c_gmax = p; p += (size_t)nb * Pp;
    s->d_out = p; p += 4 * nb;
    s->d_lam = p; p += nb;
    s->d_dl = p; p += nb + 8;
    int* q = reinterpret_cast<int*>(p);
    s->c_it = q; q += (size_t)nb * Pp;
    s->d_active = q; q += nb;
    s->d_flips = q;
    return KF_OK;
}

}  // namespace

// Solve nb ACTIVE budgets t[0..nb) at once (the caller has checked that the unconstrained minimiser violates each).
// G (possibly shifted), C: Pp-strided P x P.  K_all: nb matrices (stride Pp*Pp), solutions out; the pinned delay
// columns [fix_c0, fix_c1) must already hold their pattern in every K_b; t[] is the budget left for the free columns.
int kf_solve_l1ball_multi(kf_ctx* ctx, int P, int Pp, const double* G, const double* C, int nb, const double* t, int fix_c0,
                          int fix_c1, int max_iter, double tol, double* K_all, KfQpResult* res, cudaStream_t st) {
    if (nb <= 0) return KF_OK;
    const long long ld = Pp, kstride = (long long)Pp * Pp;
    const size_t smem = (size_t)P * (2 * sizeof(double) + sizeof(int)) + 16;
    KF_TRY(ensure_cd_smem(ctx, smem));
    Scratch sc;
    KF_TRY(carve(ctx, Pp, nb, &sc));
    KF_CUDA(ctx, ctx->d_K2.ensure((size_t)nb * kstride * sizeof(double)));   // D_b = dK_b/dlam
    double* D_all = ctx->d_K2.as<double>();

    kf_diag_kernel<<<(P + 255) / 256, 256, 0, st>>>(G, ld, P, 0.0, nullptr, sc.dG);
    kf_absmax_kernel<<<1, 1024, 0, st>>>(C, ld, P, P, fix_c0, fix_c1, sc.d_out);
    double lam_max = 0;
    KF_CUDA(ctx, cudaMemcpyAsync(&lam_max, sc.d_out, sizeof(double), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 2;
    // K_b = 0 on the free columns (K(lam_max) = 0)
    for (int b = 0; b < nb; ++b) {
        double* Kb = K_all + b * kstride;
        if (fix_c0 > 0) KF_CUDA(ctx, cudaMemset2DAsync(Kb, ld * sizeof(double), 0, (size_t)P * sizeof(double), fix_c0, st));
        if (fix_c1 < P)
            KF_CUDA(ctx, cudaMemset2DAsync(Kb + (size_t)std::max(fix_c1, 0) * ld, ld * sizeof(double), 0, (size_t)P * sizeof(double),
                                           P - std::max(fix_c1, 0), st));
    }

    // inner sweeps per evaluation are bounded so that an ill-conditioned Gram cannot run away
    const int max_sweeps = max_iter > 0 ? max_iter : (P <= 256 ? 100000 : 2000);
    const double cd_tol = tol > 0 ? tol : 1e-13;

    struct BState { double lam_hi, phi_hi, lam_lo, phi_lo, lam, l1; int side, phase, its, evals, capped; };   // phase 0 bracket, 1 secant, 2 done
    std::vector<BState> B(nb);
    std::vector<double> h_lam(nb), h_out(4 * nb), h_dl(nb, 0.0);
    std::vector<int> h_act(nb, 0);
    for (int b = 0; b < nb; ++b) B[b] = BState{lam_max, -t[b], 0, 0, lam_max, 0, 0, 0, 0, 0, 0};

    auto launch = [&](int mode, double* Kout, const double* K0) -> int {
        KF_CUDA(ctx, cudaMemcpyAsync(sc.d_lam, h_lam.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, st));
        KF_CUDA(ctx, cudaMemcpyAsync(sc.d_active, h_act.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
        CdArgs a{};
        a.G = G; a.ldg = ld; a.dG = sc.dG; a.R = C; a.ldr = ld;
        a.K = Kout; a.ldk = ld; a.K0 = K0; a.ldk0 = ld; a.kstride = kstride;
        a.P = P; a.skip0 = fix_c0; a.skip1 = fix_c1;
        a.lam = sc.d_lam; a.active = sc.d_active; a.tol = cd_tol; a.max_sweeps = max_sweeps; a.mode = mode;
        a.col_l1 = sc.c_l1; a.col_obj = sc.c_ob; a.col_aux = sc.c_aux; a.col_iters = sc.c_it;
        kf_cd_kernel<<<dim3(P, nb), CD_THREADS, smem, st>>>(a);
        kf_cd_reduce_kernel<<<nb, 256, 0, st>>>(sc.c_l1, sc.c_ob, sc.c_aux, sc.c_it, P, fix_c0, fix_c1, sc.d_active, sc.d_out);
        KF_CUDA(ctx, cudaGetLastError());
        KF_CUDA(ctx, cudaMemcpyAsync(h_out.data(), sc.d_out, sizeof(double) * 4 * nb, cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->launches += 2;
        return KF_OK;
    };

    // ---- lockstep bracketing + Illinois secant, one multiplier per budget
    for (int round = 0; round < 320; ++round) {
        int nact = 0;
        for (int b = 0; b < nb; ++b) {
            BState& s = B[b];
            h_act[b] = 0;
            if (s.phase == 2) continue;
            if (s.phase == 0) {
                s.lam *= 0.5;
                if (s.lam < 1e-300 || s.its >= 200) { s.phase = 2; continue; }
            } else {
                if (s.its >= 60 || s.lam_hi - s.lam_lo <= 1e-14 * s.lam_hi) { s.phase = 2; continue; }
                double l = (s.lam_lo * s.phi_hi - s.lam_hi * s.phi_lo) / (s.phi_hi - s.phi_lo);
                if (!(l > s.lam_lo && l < s.lam_hi)) l = 0.5 * (s.lam_lo + s.lam_hi);
                s.lam = l;
            }
            h_lam[b] = s.lam;
            h_act[b] = 1;
            ++nact;
        }
        if (!nact) break;
        KF_TRY(launch(0, K_all, nullptr));
        for (int b = 0; b < nb; ++b) {
            if (!h_act[b]) continue;
            BState& s = B[b];
            ++s.evals;
            ++s.its;
            if (h_out[4 * b + 3] >= max_sweeps) ++s.capped;
            s.l1 = h_out[4 * b + 0];
            const double phi = s.l1 - t[b];
            if (s.phase == 0) {
                if (phi > 0) { s.lam_lo = s.lam; s.phi_lo = phi; s.phase = 1; s.its = 0; s.side = 0; }
                else { s.lam_hi = s.lam; s.phi_hi = phi; }
            } else {
                if (fabs(phi) <= 1e-10 * t[b]) { s.phase = 2; continue; }
                if (phi > 0) {
                    s.lam_lo = s.lam; s.phi_lo = phi;
                    if (s.side == 1) s.phi_hi *= 0.5;
                    s.side = 1;
                } else {
                    s.lam_hi = s.lam; s.phi_hi = phi;
                    if (s.side == -1) s.phi_lo *= 0.5;
                    s.side = -1;
                }
            }
        }
    }
    // ---- exact last step on the fixed sign pattern: D = dK/dlam, ||K(lam + dl)||_1 = ||K||_1 - dl * sum s'd
    for (int b = 0; b < nb; ++b) { h_act[b] = 1; h_lam[b] = 0.0; }
    KF_TRY(launch(1, D_all, K_all));
    std::vector<int> allowed(nb, 0);
    for (int b = 0; b < nb; ++b) {
        const double den = h_out[4 * b + 2];
        if (den > 0 && B[b].l1 > 0) { h_dl[b] = (B[b].l1 - t[b]) / den; allowed[b] = 1; }
    }
    KF_CUDA(ctx, cudaMemcpyAsync(sc.d_dl, h_dl.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(sc.d_active, allowed.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemsetAsync(sc.d_flips, 0, sizeof(int) * nb, st));
    const dim3 egrid(ctx->sm_count * 2, nb);
    kf_axpy_sign_kernel<<<egrid, 256, 0, st>>>(K_all, D_all, kstride, ld, P, P, fix_c0, fix_c1, sc.d_dl, sc.d_active, sc.d_flips, 0);
    kf_axpy_sign_kernel<<<egrid, 256, 0, st>>>(K_all, D_all, kstride, ld, P, P, fix_c0, fix_c1, sc.d_dl, sc.d_active, sc.d_flips, 1);
    std::vector<int> flips(nb, 0);
    KF_CUDA(ctx, cudaMemcpyAsync(flips.data(), sc.d_flips, sizeof(int) * nb, cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 2;
    for (int b = 0; b < nb; ++b)
        if (allowed[b] && !flips[b]) B[b].lam += h_dl[b];
    // ---- evaluate; enforce feasibility to rounding
    for (int b = 0; b < nb; ++b) h_act[b] = 1;
    KF_TRY(launch(2, K_all, nullptr));
    bool rescale = false;
    std::vector<double> scl(nb, 1.0);
    for (int b = 0; b < nb; ++b)
        if (h_out[4 * b] > t[b] && h_out[4 * b] > 0) { scl[b] = t[b] / h_out[4 * b]; rescale = true; }
    if (rescale) {
        KF_CUDA(ctx, cudaMemcpyAsync(sc.d_dl, scl.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, st));
        kf_scale_kernel<<<egrid, 256, 0, st>>>(K_all, kstride, ld, P, P, fix_c0, fix_c1, sc.d_dl);
        KF_TRY(launch(2, K_all, nullptr));
    }
    for (int b = 0; b < nb; ++b) {
        res[b].l1 = h_out[4 * b + 0];
        res[b].objective = h_out[4 * b + 1];
        res[b].iters = B[b].evals;
        res[b].lam = B[b].lam;
        res[b].capped = B[b].capped;
    }
    return KF_OK;
}

// K(:, free columns) *= s   (feasibility to rounding after a solve)
// dst(:, p) = src(:, r + k R) for p = lo_r + k (gather = 1: columns dealt round-robin into the ranks' contiguous blocks), or the
// inverse dst(:, r + k R) = src(:, p) (gather = 0).  base / extra: block sizes of shard_bounds(P, r, R).
__global__ void kf_qp_deal_cols_kernel(const double* __restrict__ src, double* __restrict__ dst, long long ld, int P, int R, int gather) {
    const int p = blockIdx.x;                      // position in the dealt order
    const int base = P / R, extra = P % R;
    const int big = extra * (base + 1);
    const int r = p < big ? p / (base + 1) : extra + (p - big) / max(base, 1);
    const int k = p - (r * base + min(r, extra));
    const int c = r + k * R;                       // original column
    const double* s = src + (long long)(gather ? c : p) * ld;
    double* d = dst + (long long)(gather ? p : c) * ld;
    for (int i = threadIdx.x; i < ld; i += blockDim.x) d[i] = s[i];
}
int kf_qp_deal_cols(kf_ctx* ctx, const double* src, double* dst, int P, int Pp, int R, int gather, cudaStream_t st) {
    kf_qp_deal_cols_kernel<<<P, 256, 0, st>>>(src, dst, Pp, P, R, gather);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}

int kf_qp_scale_free(kf_ctx* ctx, double* K, int P, int Pp, int fix_c0, int fix_c1, double s, cudaStream_t st) {
    Scratch sc;
    KF_TRY(carve(ctx, Pp, 1, &sc));
    KF_CUDA(ctx, cudaMemcpyAsync(sc.d_dl, &s, sizeof(double), cudaMemcpyHostToDevice, st));
    kf_scale_kernel<<<dim3(ctx->sm_count * 2, 1), 256, 0, st>>>(K, 0, Pp, P, P, fix_c0, fix_c1, sc.d_dl);
    KF_CUDA(ctx, cudaGetLastError());
    KF_CUDA(ctx, cudaStreamSynchronize(st));     // `s` is a stack variable
    ctx->launches += 1;
    return KF_OK;
}

// G += shift * I  (the PSD conditioning branch, Ksysid.m:1119)
int kf_add_diag(kf_ctx* ctx, double* G, int Pp, int P, double shift, cudaStream_t st) {
    Scratch sc;
    KF_TRY(carve(ctx, Pp, 1, &sc));
    kf_diag_kernel<<<(P + 255) / 256, 256, 0, st>>>(G, Pp, P, shift, G, sc.dG);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}

// objective 0.5 tr(K'GK) - tr(C'K) and ||vec K||_1 over ALL P columns of one matrix, and the Frank-Wolfe duality gap
//   gap = <grad, K> + t_free * ||grad||_inf  >=  f(K) - min{ f(Z) : ||Z_free||_1 <= t_free, pinned columns fixed }
// over the free columns (grad = G K - C): a certificate of the objective that needs no reference solver.
int kf_qp_evaluate(kf_ctx* ctx, int P, int Pp, const double* G, const double* C, const double* K, int fix_c0, int fix_c1,
                   double t_free, KfQpResult* res, cudaStream_t st, int own_lo, int own_hi) {
    if (own_hi <= own_lo) { own_lo = 0; own_hi = P; }
    const long long ld = Pp;
    const size_t smem = (size_t)P * (2 * sizeof(double) + sizeof(int)) + 16;
    KF_TRY(ensure_cd_smem(ctx, smem));
    Scratch sc;
    KF_TRY(carve(ctx, Pp, 1, &sc));
    kf_diag_kernel<<<(P + 255) / 256, 256, 0, st>>>(G, ld, P, 0.0, nullptr, sc.dG);
    const double zero = 0.0;
    const int one = 1;
    KF_CUDA(ctx, cudaMemcpyAsync(sc.d_lam, &zero, sizeof(double), cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(sc.d_active, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    CdArgs a{};
    a.G = G; a.ldg = ld; a.dG = sc.dG; a.R = C; a.ldr = ld;
    a.K = const_cast<double*>(K); a.ldk = ld; a.K0 = nullptr; a.ldk0 = ld; a.kstride = 0;
    a.P = P; a.skip0 = 0; a.skip1 = 0;
    a.lam = sc.d_lam; a.active = sc.d_active; a.tol = 0; a.max_sweeps = 0; a.mode = 2;
    a.col_l1 = sc.c_l1; a.col_obj = sc.c_ob; a.col_aux = sc.c_aux; a.col_iters = sc.c_it; a.col_gmax = sc.c_gmax;
    kf_cd_kernel<<<dim3(P, 1), CD_THREADS, smem, st>>>(a);
    kf_cd_reduce_kernel<<<1, 256, 0, st>>>(sc.c_l1, sc.c_ob, sc.c_aux, sc.c_it, P, 0, 0, sc.d_active, sc.d_out);
    kf_gap_reduce_kernel<<<1, 256, 0, st>>>(sc.c_aux, sc.c_gmax, P, fix_c0, fix_c1, own_lo, own_hi, sc.d_out + 4);
    KF_CUDA(ctx, cudaGetLastError());
    double h[6];
    KF_CUDA(ctx, cudaMemcpyAsync(h, sc.d_out, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->launches += 4;
    res->l1 = h[0];
    res->objective = h[1];
    res->gap = std::max(0.0, h[4] + std::max(t_free, 0.0) * h[5]);
    res->grad_inner = h[4];
    res->grad_max = h[5];
    return KF_OK;
}
