// Fused lifting kernels: evaluate the observable dictionary on snapshot pairs
// (replaces the interpreted per-snapshot loop Ksysid.m:1030-1065 and the matlabFunction
// handles built at Ksysid.m:515,533,660,727,763,813,859).
//
// Layout: snapshots arrive column-major (one variable = one contiguous column), so a warp
// of consecutive snapshots reads and writes fully coalesced.  The panel is feature-major:
// row = observable, column = snapshot of the chunk — exactly the k-contiguous operand
// layout of the DMMA contraction in gemm.cu — and is sized to stay L2-resident, so Psi is
// produced and consumed on chip and never materialised in HBM.
//
// Parallelism: the feature program is evaluated level by level (level = dependency depth:
// variables and primitives first, then products of earlier features).  One launch per
// level; a thread owns one snapshot of one side (x = alpha / y = beta) and a slice of
// LIFT_FS features of that level, so a chunk of 1.5k snapshots x 1k features exposes
// ~4e5 threads instead of 3e3.  A product reads its two factors back from the panel
// (written by an earlier level, L2 hits).
#include "kf_internal.h"
#include "lift_eval.h"

namespace {

constexpr int LIFT_THREADS = 128;
constexpr int LIFT_FS = 8;   // features per thread

__host__ __device__ inline int pair_index(int a, int b, int m) { return a * (m + 1) - a * (a - 1) / 2 + (b - a); }

__global__ void __launch_bounds__(LIFT_THREADS) kf_lift_level_kernel(const KfLiftArgs a, int first, int count, int extras) {
    const int s = blockIdx.x * LIFT_THREADS + threadIdx.x;
    if (s >= a.Mc) return;
    const int side = blockIdx.z;   // 0: x = alpha, 1: y = beta
    const long long gs = a.start + s;
    // destination of the full dictionary: the panel section, or the scratch when dim_red follows
    double* dst = (a.n_pcs == 0) ? a.panel + (long long)(side ? a.y_off : a.x_off) * a.ld + s
                                 : a.full + (long long)side * a.n_full * a.ld + s;
    const int f0 = blockIdx.y * LIFT_FS;
    const int f1 = min(count, f0 + LIFT_FS);
    const bool valid = gs < a.M;
    const double* src = side ? a.beta : a.alpha;
    auto feat = [&](int k) -> double {
        if (k < a.nv) return k < a.nzeta ? src[(long long)k * a.M + gs] : a.u[(long long)(k - a.nzeta) * a.M + gs];
        return dst[(long long)k * a.ld];
    };
    for (int f = f0; f < f1; ++f) {
        const int j = a.order[first + f];
        double v = 0.0;   // the tail of the last chunk contributes zeros
        if (valid) {
            const KfOp op = a.ops[j];
            v = (op.kind == KF_OP_VAR) ? feat(op.a) : kf_eval_op(op, a.nv, a.centres, feat);
        }
        dst[(long long)j * a.ld] = v;
    }
    if (extras && blockIdx.y == 0 && side == 0) {
        double* sec = a.panel + (long long)a.x_off * a.ld + s;
        if (a.model == KF_LINEAR) {   // Px = [psi(x), u] (Ksysid.m:1062)
            for (int i = 0; i < a.m; ++i) sec[(long long)(a.N + i) * a.ld] = valid ? a.u[(long long)i * a.M + gs] : 0.0;
        } else if (a.model == KF_BILINEAR && a.nW > 0) {   // weights u_a*u_b of the Kronecker blocks
            double* w = a.panel + (long long)a.w_off * a.ld + s;
            for (int p = 0; p <= a.m; ++p) {
                const double up = (p && valid) ? a.u[(long long)(p - 1) * a.M + gs] : 1.0;
                for (int q = p; q <= a.m; ++q) {
                    const double uq = (q && valid) ? a.u[(long long)(q - 1) * a.M + gs] : 1.0;
                    w[(long long)pair_index(p, q, a.m) * a.ld] = valid ? (p ? KF_MUL(up, uq) : uq) : 0.0;
                }
            }
        }
    }
}

// econ lift [v; pcs' psi_full; 1] (Ksysid.m:1614-1618) from the full features in the scratch:
// thread = (snapshot, principal component slice)
__global__ void __launch_bounds__(LIFT_THREADS) kf_lift_econ_kernel(const KfLiftArgs a) {
    const int s = blockIdx.x * LIFT_THREADS + threadIdx.x;
    if (s >= a.Mc) return;
    const int side = blockIdx.z;
    const bool valid = a.start + s < a.M;
    const double* full = a.full + (long long)side * a.n_full * a.ld + s;
    double* sec = a.panel + (long long)(side ? a.y_off : a.x_off) * a.ld + s;
    const int c0 = blockIdx.y * LIFT_FS, c1 = min(a.n_pcs, c0 + LIFT_FS);
    for (int c = c0; c < c1; ++c) {
        const double* pc = a.pcs + (size_t)c * a.n_full;
        double acc = 0.0;
        for (int j = 0; j < a.n_full; ++j) acc = fma(pc[j], full[(long long)j * a.ld], acc);
        sec[(long long)(a.nv + c) * a.ld] = valid ? acc : 0.0;
    }
    if (blockIdx.y == 0) {
        for (int i = 0; i < a.nv; ++i) sec[(long long)i * a.ld] = full[(long long)i * a.ld];
        sec[(long long)(a.nv + a.n_pcs) * a.ld] = valid ? 1.0 : 0.0;
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

// Dependency levels of the program: level 0 = variables and primitives (functions of v only),
// level L = products whose deepest factor has level L-1.  `order` lists the features level by level.
void kf_program_levels(const KfProgram& p, std::vector<int>& order, std::vector<int>& level_start) {
    const int n = p.n_full();
    std::vector<int> depth(n, 0);
    int maxd = 0;
    for (int j = 0; j < n; ++j) {
        const KfOp& op = p.ops[j];
        if (op.kind == KF_OP_MUL) {
            const int da = op.a < p.nv ? -1 : depth[op.a], db = op.b < p.nv ? -1 : depth[op.b];
            depth[j] = 1 + std::max(da, db);
        }
        maxd = std::max(maxd, depth[j]);
    }
    order.clear();
    level_start.assign(1, 0);
    for (int d = 0; d <= maxd; ++d) {
        for (int j = 0; j < n; ++j)
            if (depth[j] == d) order.push_back(j);
        level_start.push_back((int)order.size());
    }
}

int kf_launch_lift(kf_ctx* ctx, const KfLiftArgs& a, cudaStream_t st) {
    const int nsides = a.nsides > 0 ? a.nsides : 2;
    const unsigned gx = (unsigned)((a.Mc + LIFT_THREADS - 1) / LIFT_THREADS);
    const int nlev = (int)ctx->level_start.size() - 1;
    for (int l = 0; l < nlev; ++l) {
        const int first = ctx->level_start[l], count = ctx->level_start[l + 1] - first;
        if (count <= 0) continue;
        dim3 grid(gx, (count + LIFT_FS - 1) / LIFT_FS, nsides);
        kf_lift_level_kernel<<<grid, LIFT_THREADS, 0, st>>>(a, first, count, (l == 0 && a.extras) ? 1 : 0);
        ctx->launches += 1;
    }
    if (a.n_pcs > 0) {
        dim3 grid(gx, (a.n_pcs + LIFT_FS - 1) / LIFT_FS, nsides);
        kf_lift_econ_kernel<<<grid, LIFT_THREADS, 0, st>>>(a);
        ctx->launches += 1;
    }
    KF_CUDA(ctx, cudaGetLastError());
    return KF_OK;
}

int kf_launch_lift_points(kf_ctx* ctx, const KfOp* ops, const double* centres, const double* pcs, int nv, int n_full,
                          int n_pcs, const double* V, long long rows, double* full, double* out, long long ldo,
                          cudaStream_t st) {
    if (rows <= 0) return KF_OK;
    KfLiftArgs a{};
    a.ops = ops; a.centres = centres; a.pcs = pcs; a.order = ctx->d_order.as<int>();
    a.nv = nv; a.n_full = n_full; a.n_pcs = n_pcs; a.N = n_pcs ? nv + n_pcs + 1 : n_full;
    a.nzeta = nv; a.m = 0; a.model = KF_NONLINEAR;
    a.alpha = V; a.beta = V; a.u = V;
    a.M = rows; a.start = 0; a.Mc = (int)rows;
    a.panel = out; a.ld = ldo; a.full = full;
    a.x_off = 0; a.y_off = 0; a.w_off = 0; a.nW = 0;
    a.nsides = 1; a.extras = 0;
    if (n_pcs && ldo != rows) {
        ctx->err = "kf_launch_lift_points: dim_red needs ldo == rows";
        return KF_EINVAL;
    }
    return kf_launch_lift(ctx, a, st);
}

int kf_launch_regressors(kf_ctx* ctx, const KfLiftArgs& a0, double* AB, double* /*unused*/, long long ldp,
                         cudaStream_t st) {
    // psi(x) -> columns [0,N), psi(y) -> columns [P, P+N) of AB (ld = ldp >= M); whole data set at once
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
    a.nsides = 2;
    a.extras = 0;
    KF_TRY(kf_launch_lift(ctx, a, st));
    if (a.model != KF_NONLINEAR) {
        dim3 grid((unsigned)((a.M + 127) / 128), a.model == KF_BILINEAR ? (a.N + 63) / 64 : 1);
        kf_regressor_post_kernel<<<grid, 128, 0, st>>>(a.model, a.N, P, a.m, a.u, a.M, AB, ldp);
        KF_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    return KF_OK;
}
