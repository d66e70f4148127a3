// Fused lifting kernels: evaluate the observable dictionary on snapshot pairs
// (replaces the interpreted per-snapshot loop Ksysid.m:1030-1065 and the matlabFunction
// handles built at Ksysid.m:515,533,660,727,763,813,859).
//
// Layout: snapshots arrive column-major (one variable = one contiguous column), so a
// warp of consecutive snapshots reads and writes fully coalesced.  One thread owns one
// snapshot of one side (x = alpha, y = beta) and walks the feature program in index
// order; a MUL op reads two earlier features of the same snapshot back from the panel
// (same-thread RAW through L1/L2 — the panel never has to leave L2).  The panel is
// feature-major: row = observable, column = snapshot of the chunk, which is exactly the
// k-contiguous operand layout of the DMMA contraction in gemm.cu.
#include "kf_internal.h"
#include "lift_eval.h"

namespace {

constexpr int LIFT_THREADS = 128;

template <class VarFn>
__device__ __forceinline__ void lift_one(const KfOp* __restrict__ ops, const double* __restrict__ centres, int nv,
                                         int n_full, VarFn var, double* F, long long ld) {
    for (int i = 0; i < nv; ++i) F[(long long)i * ld] = var(i);
    for (int j = nv; j < n_full; ++j) {
        const KfOp op = ops[j];
        F[(long long)j * ld] = kf_eval_op(op, nv, centres, [&](int k) { return F[(long long)k * ld]; });
    }
}

// econ lift [v; pcs' psi_full; 1] (Ksysid.m:1614-1618) from the full features in `full`
__device__ __forceinline__ void econ_rows(const double* __restrict__ pcs, int nv, int n_full, int n_pcs,
                                          const double* full, long long ldf, double* out, long long ldo) {
    for (int i = 0; i < nv; ++i) out[(long long)i * ldo] = full[(long long)i * ldf];
    for (int c = 0; c < n_pcs; ++c) {
        const double* pc = pcs + (size_t)c * n_full;
        double acc = 0.0;
        for (int j = 0; j < n_full; ++j) acc = fma(pc[j], full[(long long)j * ldf], acc);
        out[(long long)(nv + c) * ldo] = acc;
    }
    out[(long long)(nv + n_pcs) * ldo] = 1.0;
}

// weight row index of the pair (a,b), a<=b, in 0..m (u_0 = 1)
__host__ __device__ inline int pair_index(int a, int b, int m) { return a * (m + 1) - a * (a - 1) / 2 + (b - a); }

__global__ void __launch_bounds__(LIFT_THREADS) kf_lift_panel_kernel(const KfLiftArgs a) {
    const int s = blockIdx.x * LIFT_THREADS + threadIdx.x;
    if (s >= a.Mc) return;
    const int side = blockIdx.y;   // 0: x = alpha, 1: y = beta
    const long long gs = a.start + s;
    double* sec = a.panel + (long long)(side ? a.y_off : a.x_off) * a.ld + s;
    const int sec_rows = a.N + ((side == 0 && a.model == KF_LINEAR) ? a.m : 0);
    if (gs >= a.M) {   // tail of the last chunk contributes zeros
        for (int j = 0; j < sec_rows; ++j) sec[(long long)j * a.ld] = 0.0;
        if (side == 0)
            for (int q = 0; q < a.nW; ++q) a.panel[(long long)(a.w_off + q) * a.ld + s] = 0.0;
        return;
    }
    const double* src = side ? a.beta : a.alpha;
    auto var = [&](int i) -> double {
        return i < a.nzeta ? src[(long long)i * a.M + gs] : a.u[(long long)(i - a.nzeta) * a.M + gs];
    };
    if (a.n_pcs == 0) {
        lift_one(a.ops, a.centres, a.nv, a.n_full, var, sec, a.ld);
    } else {
        double* full = a.full + (long long)side * a.n_full * a.ld + s;
        lift_one(a.ops, a.centres, a.nv, a.n_full, var, full, a.ld);
        econ_rows(a.pcs, a.nv, a.n_full, a.n_pcs, full, a.ld, sec, a.ld);
    }
    if (side == 0) {
        if (a.model == KF_LINEAR) {   // Px = [psi(x), u] (Ksysid.m:1062)
            for (int i = 0; i < a.m; ++i) sec[(long long)(a.N + i) * a.ld] = a.u[(long long)i * a.M + gs];
        } else if (a.model == KF_BILINEAR) {   // weights u_a*u_b of the Kronecker blocks
            double* w = a.panel + (long long)a.w_off * a.ld + s;
            for (int p = 0; p <= a.m; ++p) {
                const double up = p ? a.u[(long long)(p - 1) * a.M + gs] : 1.0;
                for (int q = p; q <= a.m; ++q) {
                    const double uq = q ? a.u[(long long)(q - 1) * a.M + gs] : 1.0;
                    w[(long long)pair_index(p, q, a.m) * a.ld] = p ? KF_MUL(up, uq) : uq;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(LIFT_THREADS)
kf_lift_points_kernel(const KfOp* __restrict__ ops, const double* __restrict__ centres, const double* __restrict__ pcs,
                      int nv, int n_full, int n_pcs, const double* __restrict__ V, long long rows, double* full,
                      double* out, long long ldo) {
    const long long s = (long long)blockIdx.x * LIFT_THREADS + threadIdx.x;
    if (s >= rows) return;
    auto var = [&](int i) -> double { return V[(long long)i * rows + s]; };
    if (n_pcs == 0) {
        lift_one(ops, centres, nv, n_full, var, out + s, ldo);
    } else {
        lift_one(ops, centres, nv, n_full, var, full + s, rows);
        econ_rows(pcs, nv, n_full, n_pcs, full + s, rows, out + s, ldo);
    }
}

// Complete materialised regressors AB = [Px | Py] (M x 2P, ld): columns [0,N) and [P,P+N)
// already hold psi(x), psi(y).  linear: append u to both (Ksysid.m:1062-1063); bilinear:
// blocks u_k * psi (Ksysid.m:510-511).
__global__ void kf_regressor_post_kernel(int model, int N, int P, int m, const double* __restrict__ u, long long M,
                                         double* AB, long long ld) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= M) return;
    if (model == KF_LINEAR) {
        for (int i = 0; i < m; ++i) {
            const double ui = u[(long long)i * M + s];
            AB[(long long)(N + i) * ld + s] = ui;
            AB[(long long)(P + N + i) * ld + s] = ui;
        }
    } else if (model == KF_BILINEAR) {
        const int j0 = blockIdx.y * 64;
        for (int k = 0; k < m; ++k) {
            const double uk = u[(long long)k * M + s];
            for (int j = j0; j < min(N, j0 + 64); ++j) {
                AB[(long long)((k + 1) * N + j) * ld + s] = KF_MUL(uk, AB[(long long)j * ld + s]);
                AB[(long long)(P + (k + 1) * N + j) * ld + s] = KF_MUL(uk, AB[(long long)(P + j) * ld + s]);
            }
        }
    }
}

}  // namespace

int kf_launch_lift(kf_ctx* ctx, const KfLiftArgs& a, cudaStream_t st) {
    dim3 grid((a.Mc + LIFT_THREADS - 1) / LIFT_THREADS, 2);
    kf_lift_panel_kernel<<<grid, LIFT_THREADS, 0, st>>>(a);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}

int kf_launch_lift_points(kf_ctx* ctx, const KfOp* ops, const double* centres, const double* pcs, int nv, int n_full,
                          int n_pcs, const double* V, long long rows, double* full, double* out, long long ldo,
                          cudaStream_t st) {
    if (rows <= 0) return KF_OK;
    unsigned grid = (unsigned)((rows + LIFT_THREADS - 1) / LIFT_THREADS);
    kf_lift_points_kernel<<<grid, LIFT_THREADS, 0, st>>>(ops, centres, pcs, nv, n_full, n_pcs, V, rows, full, out, ldo);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}

int kf_launch_regressors(kf_ctx* ctx, const KfLiftArgs& a0, double* AB, double* /*unused*/, long long ldp,
                         cudaStream_t st) {
    // psi(x) -> columns [0,N), psi(y) -> columns [P, P+N) of AB (ld = ldp >= M); whole data set in one launch
    KfLiftArgs a = a0;
    const int P = kf_regressor_width(a.model, a.N, a.m);
    a.panel = AB;
    a.ld = ldp;
    a.start = 0;
    a.Mc = (int)a.M;
    a.x_off = 0;
    a.y_off = P;
    a.nW = 0;
    a.w_off = 0;
    const int model = a.model;
    if (model == KF_BILINEAR) a.model = KF_NONLINEAR + 100;   // suppress the weight rows / u rows of panel mode
    if (model == KF_LINEAR) a.model = KF_NONLINEAR + 100;
    KF_TRY(kf_launch_lift(ctx, a, st));
    if (model != KF_NONLINEAR) {
        dim3 grid((unsigned)((a.M + 127) / 128), model == KF_BILINEAR ? (a.N + 63) / 64 : 1);
        kf_regressor_post_kernel<<<grid, 128, 0, st>>>(model, a.N, P, a.m, a.u, a.M, AB, ldp);
        KF_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    return KF_OK;
}
