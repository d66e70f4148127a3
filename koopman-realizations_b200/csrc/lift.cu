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
#include <algorithm>
#include <cstring>

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
// already hold psi(x), psi(y).  Kronecker blocks u_ka * (w_kc * psi) (bilinear: Ksysid.m:510-511; loaded: 594-599, 604-605),
// and for the linear model u appended to both (Ksysid.m:1062-1063).
__global__ void kf_regressor_post_kernel(int model, int N, int P, int m, int nw, const double* __restrict__ u, const double* __restrict__ w,
                                         long long M, long long ldu, double* AB, long long ld) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= M) return;
    const int nku = model == KF_BILINEAR ? m : 0;
    const int j0 = blockIdx.y * 64;
    for (int ka = 0; ka <= nku; ++ka)
        for (int kc = 0; kc <= nw; ++kc) {
            if (ka == 0 && kc == 0) continue;
            const double ua = ka ? u[(long long)(ka - 1) * ldu + s] : 1.0, wc = kc ? w[(long long)(kc - 1) * ldu + s] : 1.0;
            const long long col = (long long)(ka * (nw + 1) + kc) * N;
            for (int j = j0; j < min(N, j0 + 64); ++j) {
                double x = AB[(long long)j * ld + s], y = AB[(long long)(P + j) * ld + s];
                if (kc) { x = KF_MUL(wc, x); y = KF_MUL(wc, y); }
                if (ka) { x = KF_MUL(ua, x); y = KF_MUL(ua, y); }
                AB[(col + j) * ld + s] = x;
                AB[(P + col + j) * ld + s] = y;
            }
        }
    if (model == KF_LINEAR && blockIdx.y == 0) {
        for (int i = 0; i < m; ++i) {
            const double ui = u[(long long)i * ldu + s];
            AB[((long long)N * (nw + 1) + i) * ld + s] = ui;
            AB[((long long)P + (long long)N * (nw + 1) + i) * ld + s] = ui;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Materialising lift (lift-only mode: kf_lift / kf_lift_dev, koopData.Px / Py, the QRCP route).  HBM-bound by
// construction — 8 (2 nzeta + m) B read and 16 P B written per pair — so every byte must move once:
// a CTA evaluates the WHOLE feature program for a tile of LT_S snapshots with all features in shared memory
// (products read their factors from shared memory, not back from HBM as the level kernel would for an
// HBM-resident output), then streams the rows of [Px | Py] (or Psi) out once, 128-byte segments per feature;
// the bilinear blocks u_k psi and the dim_red projection are formed from shared memory too.
constexpr int LT_THREADS = 256;
constexpr int LT_MAXLEV = KF_LT_MAXLEV;
constexpr int LT_SLOTS = 440;      // target size of a feature group (shared-memory slots): 3 CTAs per SM with 16-snapshot tiles

// A feature GROUP (LtGroup / LtOp, kf_internal.h): a contiguous range of output features plus everything they depend on
// (their dependency closure), renumbered into compact shared-memory slots.  Splitting a large dictionary into groups that
// run as separate CTAs keeps the shared memory per CTA small (several CTAs per SM) at the price of re-evaluating shared
// parents (cheap products).
struct KfLiftTileArgs {
    const KfOp* ops; const double* centres; const double* pcs; const int* order;   // ops / order: unused by the tile kernel
    const LtOp* gops; const LtStore* gstore; const LtGroup* groups; int ngroups;
    int nv, n_full, n_pcs, N;          // N = lifted dimension (n_full, or nv + n_pcs + 1 with dim_red)
    int nzeta, m, model;
    int mode;                           // 0: points V -> Psi (rows x N);  1: regressors [Px | Py] (M x 2P)
    int P;
    int side_off;                       // mode 0 with two sides (panel mode): psi(beta) goes to rows side_off .. of `out`
    const unsigned long long* rops;     // streaming kernel: packed row op of every output row (LtRop)
    const double* rconst;               // ... constants of the primitive row ops
    int ncent;                          // ... centre doubles staged in shared memory (0: read through L1)
    int max_slots, max_ops, max_nst;
    const double* alpha; const double* beta; const double* u; long long M;   // M = points of this launch
    const double* w; int nw;            // loads of a `loaded` model (mode 1): blocks w_c psi, Ksysid.m:594-599
    long long ldin;                     // leading dimension of alpha / beta / u (>= M; a chunk of a longer column)
    double* out; long long ld;
};

__device__ __forceinline__ void lt_ld4(const double* p, double (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void lt_st4(double* p, const double (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

// LS = snapshots per tile (8 or 16: 64- or 128-byte segments per output row), chosen by the host so that several CTAs fit
// an SM and the load / evaluate / store phases of different tiles overlap.  blockIdx.y = side * ngroups + group.
// Work split: evaluation — a thread takes one op (read from the shared-memory copy of the group's level-ordered program)
// and applies it to 4 snapshots with 128-bit shared-memory accesses, so the interpretive overhead is paid once per
// 4 elements; store — 2 snapshots (one 128-bit store) per thread.
template <int LS>
__global__ void __launch_bounds__(LT_THREADS) kf_lift_tile_kernel(const KfLiftTileArgs a) {
    extern __shared__ __align__(16) double lt_smem[];
    double* sh = lt_smem;                              // [max_slots][LS]; slots 0 .. nv-1 are the variables
    double* su = sh + (size_t)a.max_slots * LS;        // [m + nw][LS]   inputs u, then loads w, of the tile
    double* se = su + (size_t)(a.m + a.nw) * LS;       // [N][LS]   econ features (dim_red only, single group)
    LtOp* sop = reinterpret_cast<LtOp*>(se + (a.n_pcs > 0 ? (size_t)a.N * LS : 0));   // [max_ops] in level order
    LtStore* sst = reinterpret_cast<LtStore*>(sop + a.max_ops);                       // [max_nst] (slot, output row)
    __shared__ LtGroup grp;
    const int tid = threadIdx.x;
    const int side = blockIdx.y / a.ngroups, g = blockIdx.y % a.ngroups;
    const double* src = side ? a.beta : a.alpha;
    const long long ntiles = (a.M + LS - 1) / LS;
    const bool vec2 = ((a.ld & 1) == 0) && ((reinterpret_cast<unsigned long long>(a.out) & 15ull) == 0);
    if (tid == 0) grp = a.groups[g];
    __syncthreads();
    for (int e = tid; e < grp.nops; e += LT_THREADS) sop[e] = a.gops[grp.op_off + e];
    for (int e = tid; e < grp.nst; e += LT_THREADS) sst[e] = a.gstore[grp.st_off + e];
    __syncthreads();
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long g0 = tile * LS;
        // ---- A: the variables v (slots 0 .. nv-1) and u of the tile; the tail of the last tile is zero
        for (int idx = tid; idx < (a.nv + a.m + a.nw) * LS; idx += LT_THREADS) {
            const int k = idx / LS, s = idx & (LS - 1);
            const long long gs = g0 + s;
            double v = 0.0;
            if (gs < a.M) {
                if (k < a.nv) v = k < a.nzeta ? src[(long long)k * a.ldin + gs] : a.u[(long long)(k - a.nzeta) * a.ldin + gs];
                else if (k < a.nv + a.m) v = a.u[(long long)(k - a.nv) * a.ldin + gs];
                else v = a.w[(long long)(k - a.nv - a.m) * a.ldin + gs];
            }
            if (k < a.nv) sh[k * LS + s] = v; else su[(k - a.nv) * LS + s] = v;
        }
        __syncthreads();
        // ---- B: the group's program, level by level
        {
            constexpr int QS = LS / 4;                 // snapshot quads per tile
            constexpr int NF = LT_THREADS / QS;
            const int s0 = (tid % QS) * 4, fs = tid / QS;
            for (int l = 0; l < grp.nlevels; ++l) {
                const int first = grp.level_start[l], last = grp.level_start[l + 1];
                for (int e = first + fs; e < last; e += NF) {
                    const LtOp op = sop[e];
                    double v[4];
                    if (op.kind == KF_OP_MUL) {
                        double x[4], y[4];
                        lt_ld4(sh + op.a * LS + s0, x);
                        lt_ld4(sh + op.b * LS + s0, y);
#pragma unroll
                        for (int t = 0; t < 4; ++t) v[t] = KF_MUL(x[t], y[t]);
                    } else if (op.kind == KF_OP_VAR) {
                        lt_ld4(sh + op.a * LS + s0, v);
                    } else {
                        KfOp o{};
                        o.kind = op.kind; o.a = op.a; o.b = op.b; o.c = op.c;
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const double* col = sh + s0 + t;
                            v[t] = kf_eval_op(o, a.nv, a.centres, [&](int k) -> double { return col[k * LS]; });
                        }
                    }
                    lt_st4(sh + op.j * LS + s0, v);
                }
                __syncthreads();
            }
        }
        if (a.n_pcs > 0) {      // econ lift [v; pcs' psi_full; 1]  (Ksysid.m:1614-1618); single group, slot = feature index
            const int s = tid & (LS - 1), fs = tid / LS;
            constexpr int NF = LT_THREADS / LS;
            const bool valid = g0 + s < a.M;
            for (int c = fs; c < a.n_pcs; c += NF) {
                const double* pc = a.pcs + (size_t)c * a.n_full;
                double acc = 0.0;
                for (int j = 0; j < a.n_full; ++j) acc = fma(pc[j], sh[j * LS + s], acc);
                se[(a.nv + c) * LS + s] = valid ? acc : 0.0;
            }
            for (int i = fs; i < a.nv; i += NF) se[i * LS + s] = sh[i * LS + s];
            if (fs == 0) se[(a.nv + a.n_pcs) * LS + s] = valid ? 1.0 : 0.0;
            __syncthreads();
        }
        // ---- C: stream the stored rows out, once
        {
            constexpr int PS = LS / 2;                 // snapshot pairs per tile
            constexpr int NR = LT_THREADS / PS;
            const int p2 = (tid % PS) * 2, rr = tid / PS;
            const long long gs = g0 + p2;
            const bool ok0 = gs < a.M, ok1 = gs + 1 < a.M;
            auto put = [&](double* dst, double v0, double v1) {
                if (ok1 && vec2) *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                else {
                    if (ok0) dst[0] = v0;
                    if (ok1) dst[1] = v1;
                }
            };
            if (ok0) {
                double* base = a.out + (long long)(side ? (a.mode == 0 ? a.side_off : a.P) : 0) * a.ld + gs;
                const bool direct = a.n_pcs > 0 || a.ngroups == 1;       // slot = output row = feature index
                const int nrow = direct ? a.N : grp.nst;
                const double* psi = a.n_pcs > 0 ? se : sh;
                for (int e = rr; e < nrow; e += NR) {
                    int slot = e, row = e;
                    if (!direct) { const LtStore sr = sst[e]; slot = sr.slot; row = sr.row; }
                    const double2 v = *reinterpret_cast<const double2*>(psi + slot * LS + p2);
                    put(base + (long long)row * a.ld, v.x, v.y);
                    if (a.mode == 1 && (a.model == KF_BILINEAR || a.nw > 0)) {
                        // Kronecker blocks: column (ka (nw+1) + kc) N + row holds u_ka * (w_kc * psi)  — [1; u] (x) [1; w] (x) psi with
                        // the reference's association (psi_L = [psi; w_c psi] first, Ksysid.m:594-599; then u_k psi_L, 510-511 / 604-605)
                        const int nku = a.model == KF_BILINEAR ? a.m : 0;
                        for (int ka = 0; ka <= nku; ++ka)
                            for (int kc = 0; kc <= a.nw; ++kc) {
                                if (ka == 0 && kc == 0) continue;
                                double x0 = v.x, x1 = v.y;
                                if (kc) { x0 = KF_MUL(su[(a.m + kc - 1) * LS + p2], x0); x1 = KF_MUL(su[(a.m + kc - 1) * LS + p2 + 1], x1); }
                                if (ka) { x0 = KF_MUL(su[(ka - 1) * LS + p2], x0); x1 = KF_MUL(su[(ka - 1) * LS + p2 + 1], x1); }
                                put(base + ((long long)(ka * (a.nw + 1) + kc) * a.N + row) * a.ld, x0, x1);
                            }
                    }
                }
                if (a.mode == 1 && a.model == KF_LINEAR && g == 0) {   // [psi_L, u]  (Ksysid.m:1062-1063)
                    for (int i = rr; i < a.m; i += NR)
                        put(base + ((long long)a.N * (a.nw + 1) + i) * a.ld, su[i * LS + p2], su[i * LS + p2 + 1]);
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming variant (the default for n_pcs == 0).  Measured on this pool's B200 (tools/store_pattern_bench.cu,
// profiles/r02_store_pattern_microbench.txt; profiles/r02_lift_stream_ncu_summary.txt):
//  (1) what the DRAM sees decides the write rate, not the store mechanism: a pure store kernel with this access pattern
//      (feature-major rows, one tile of LS snapshots at a time) writes 3.7-4.2 TB/s with 128-byte runs, 4.6 / 5.9 TB/s with
//      256-byte runs on unaligned / 128-byte-aligned rows, 5.0 / 5.9 TB/s with 512-byte runs — st.global and TMA tensor stores
//      (UTMASTG) give the same figures;
//  (2) kf_lift_tile_kernel spends its issue slots, not bandwidth: ncu's source view put 60 % of its 1.0e9 warp instructions
//      into the table-driven store loop and 30 % into per-element gaussian evaluation, 1.6 IPC, 25 % of the warps resident.
// Hence: tiles of 32 (aligned rows) or 64 snapshots; only the features that other features READ (for a monomial dictionary
// of degree d: the degrees < d) are held in shared memory; every output row is produced by ONE row op in registers — a product
// of two slots, a gaussian of the variables, a copy of a slot — applied to 4 snapshots per thread and stored straight to
// global memory as complete half runs, together with its bilinear blocks u_k psi (u_k in registers).  Leaves never touch shared
// memory, the store phase has no table lookup, no barrier and ~10 instructions per lifted element, and since the slots are
// few a whole dictionary usually is ONE group: the inputs of a tile are read once and nothing is evaluated twice.  The next
// tile's inputs are fetched into registers before the row phase, so their latency hides under the stores.
// a row op packed into 8 bytes (shared memory: one per output row of the group)
struct LtRop { unsigned short a, b; unsigned char kind, pad; unsigned short ci; };
static_assert(sizeof(LtRop) == 8, "LtRop is 8 bytes");

// a thread's 4 snapshots of a row: the pairs 2t, 2t+1 and LS/2 + 2t, LS/2 + 2t + 1 — every 128-bit access of LS/4 consecutive
// threads covers a contiguous half run (128 B at LS = 32, 256 B at LS = 64)
template <int LS> __device__ __forceinline__ void lq_ld(const double* row, int t2, double (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(row + t2), b = *reinterpret_cast<const double2*>(row + LS / 2 + t2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <int LS> __device__ __forceinline__ void lq_st(double* row, int t2, const double (&v)[4]) {
    *reinterpret_cast<double2*>(row + t2) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(row + LS / 2 + t2) = make_double2(v[2], v[3]);
}

// one op on a snapshot quad; operands are shared-memory slots (GAUSS: the variables, slots 0 .. nv-1; centres in shared memory
// or read through L1).  c is only read by the primitive kinds.
template <int LS, class GetC>
__device__ __forceinline__ void lq_eval(int kind, int oa, int ob, GetC getc, const double* sh, int t2, int nv, const double* cen, double (&v)[4]) {
    if (kind == KF_OP_MUL) {
        double x[4], y[4];
        lq_ld<LS>(sh + oa * LS, t2, x);
        lq_ld<LS>(sh + ob * LS, t2, y);
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = KF_MUL(x[t], y[t]);
    } else if (kind == KF_OP_GAUSS) {                    // exp(-||v - c||^2): the operation order of kf_eval_op, the centre read once per quad
        const double* c = cen + (size_t)oa * nv;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < nv; ++i) {
            const double ci = c[i];
            double x[4];
            lq_ld<LS>(sh + i * LS, t2, x);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const double d = KF_SUB(x[t], ci);
                acc[t] = KF_ADD(acc[t], KF_MUL(d, d));
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = exp(-acc[t]);
    } else if (kind == KF_OP_VAR) {
        lq_ld<LS>(sh + oa * LS, t2, v);
    } else {
        KfOp o{};
        o.kind = kind; o.a = oa; o.b = ob; o.c = getc();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const double* col = sh + (t < 2 ? t2 + t : LS / 2 + t2 + t - 2);
            v[t] = kf_eval_op(o, nv, cen, [&](int k) -> double { return col[k * LS]; });
        }
    }
}

__device__ __forceinline__ void lw_cp_async8(double* dst_smem, const double* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 8 : 0;                         // 0 source bytes: the destination is zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(d), "l"(src), "r"(n) : "memory");
}

template <int LS, int MINB>
__global__ void __launch_bounds__(LT_THREADS, MINB) kf_lift_stream_kernel(const KfLiftTileArgs a) {
    extern __shared__ __align__(16) double lt_smem[];
    const int nrows_in = a.nv + a.m + a.nw;
    double* sh = lt_smem;                              // [max_slots][LS]; slots 0 .. nv-1 are the variables
    double* su = sh + (size_t)a.max_slots * LS;        // [m + nw][LS]   inputs u, then loads w, of the tile
    double* stage = su + (size_t)(a.m + a.nw) * LS;    // [nv + m + nw][LS]  landing buffer of the NEXT tile's inputs (cp.async)
    double* scen = stage + (size_t)nrows_in * LS;      // [ncent] gaussian centres (when they fit)
    LtOp* sop = reinterpret_cast<LtOp*>(scen + a.ncent);                        // [max_ops] slot ops in level order
    LtRop* srop = reinterpret_cast<LtRop*>(sop + a.max_ops);                    // [max_nst] packed row ops
    const double** inrow = reinterpret_cast<const double**>(srop + a.max_nst);  // [nv + m + nw] first element of every input row
    __shared__ LtGroup grp;
    const int tid = threadIdx.x;
    const int side = blockIdx.y / a.ngroups, g = blockIdx.y % a.ngroups;
    const long long ntiles = (a.M + LS - 1) / LS;
    if (tid == 0) grp = a.groups[g];
    for (int k = tid; k < nrows_in; k += LT_THREADS) {
        const double* r;
        if (k < a.nv) r = k < a.nzeta ? (side ? a.beta : a.alpha) + (long long)k * a.ldin : a.u + (long long)(k - a.nzeta) * a.ldin;
        else if (k < a.nv + a.m) r = a.u + (long long)(k - a.nv) * a.ldin;
        else r = a.w + (long long)(k - a.nv - a.m) * a.ldin;
        inrow[k] = r;
    }
    __syncthreads();
    for (int e = tid; e < grp.nops; e += LT_THREADS) sop[e] = a.gops[grp.op_off + e];
    for (int e = tid; e < grp.nst; e += LT_THREADS) reinterpret_cast<unsigned long long*>(srop)[e] = a.rops[grp.row0 + e];
    for (int e = tid; e < a.ncent; e += LT_THREADS) scen[e] = a.centres[e];
    const double* cen = a.ncent > 0 ? scen : a.centres;
    // inputs of a tile: (row k, snapshot sn) -> stage[k][sn], asynchronously; the tail of the last tile is zero-filled
    constexpr int KSTEP = LT_THREADS / LS;
    const int sn = tid & (LS - 1), k0 = tid / LS;
    auto fetch_tile = [&](long long g0) {
        const long long gs = g0 + sn;
        const bool ok = gs < a.M;
        for (int k = k0; k < nrows_in; k += KSTEP) lw_cp_async8(stage + k * LS + sn, inrow[k] + (ok ? gs : 0), ok);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch_tile((long long)blockIdx.x * LS);
    constexpr int QS = LS / 4, NF = LT_THREADS / QS;   // threads per row, rows in flight per CTA
    const int t2 = (tid % QS) * 2, fs = tid / QS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long g0 = tile * LS;
        // ---- A: the inputs that landed in the stage buffer -> variable slots and u / w rows
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        for (int k = k0; k < nrows_in; k += KSTEP) (k < a.nv ? sh + k * LS : su + (k - a.nv) * LS)[sn] = stage[k * LS + sn];
        __syncthreads();
        // the next tile of this CTA: in flight during the evaluation and the row phase
        if (tile + gridDim.x < ntiles) fetch_tile((tile + gridDim.x) * LS);
        // ---- B: the features other features read, level by level, into their slots
        for (int l = 0; l < grp.nlevels; ++l) {
            const int first = grp.level_start[l], last = grp.level_start[l + 1];
            for (int e = first + fs; e < last; e += NF) {
                const LtOp op = sop[e];
                double v[4];
                lq_eval<LS>(op.kind, op.a, op.b, [&]() { return op.c; }, sh, t2, a.nv, cen, v);
                lq_st<LS>(sh + op.j * LS, t2, v);
            }
            if (last > first) __syncthreads();
        }
        // ---- C: one row op per output row, evaluated in registers and stored with its blocks u_ka (w_kc psi)
        {
            const long long gA = g0 + t2, gB = gA + LS / 2;
            const bool full = g0 + LS <= a.M;
            const bool okA0 = gA < a.M, okA1 = gA + 1 < a.M, okB0 = gB < a.M, okB1 = gB + 1 < a.M;
            auto put = [&](double* dst, const double (&x)[4]) {      // dst: the row's element of snapshot gA (16-byte aligned: launcher)
                if (full) {
                    *reinterpret_cast<double2*>(dst) = make_double2(x[0], x[1]);
                    *reinterpret_cast<double2*>(dst + LS / 2) = make_double2(x[2], x[3]);
                } else {
                    if (okA1) *reinterpret_cast<double2*>(dst) = make_double2(x[0], x[1]); else if (okA0) dst[0] = x[0];
                    if (okB1) *reinterpret_cast<double2*>(dst + LS / 2) = make_double2(x[2], x[3]); else if (okB0) dst[LS / 2] = x[2];
                }
            };
            auto row = [&](int e, double (&v)[4]) {
                const LtRop r = srop[e];
                lq_eval<LS>(r.kind, r.a, r.b, [&]() { return __ldg(a.rconst + r.ci); }, sh, t2, a.nv, cen, v);
            };
            double* base = a.out + (long long)(side ? (a.mode == 0 ? a.side_off : a.P) : 0) * a.ld + gA;
            const long long rstep = (long long)NF * a.ld, bstep = (long long)a.N * a.ld;
            double* dst = base + (long long)(grp.row0 + fs) * a.ld;
            const int nku = (a.mode == 1 && a.model == KF_BILINEAR) ? a.m : 0;
            const int nw = a.mode == 1 ? a.nw : 0;
            if (nw == 0 && nku == 0) {
                for (int e = fs; e < grp.nst; e += NF, dst += rstep) {
                    double v[4];
                    row(e, v);
                    put(dst, v);
                }
            } else if (nw == 0 && nku <= 3) {                // u_k psi, the u_k quads in registers  (Ksysid.m:510-511)
                double uu[3][4];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (k < nku) lq_ld<LS>(su + k * LS, t2, uu[k]);
                    else { uu[k][0] = uu[k][1] = uu[k][2] = uu[k][3] = 0.0; }
                }
                for (int e = fs; e < grp.nst; e += NF, dst += rstep) {
                    double v[4];
                    row(e, v);
                    put(dst, v);
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (k < nku) {
                            double x[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t) x[t] = KF_MUL(uu[k][t], v[t]);
                            put(dst + (k + 1) * bstep, x);
                        }
                }
            } else {
                // column (ka (nw+1) + kc) N + row holds u_ka * (w_kc * psi) — [1; u] (x) [1; w] (x) psi with the reference's
                // association (psi_L = [psi; w_c psi] first, Ksysid.m:594-599; then u_k psi_L, 510-511 / 604-605)
                for (int e = fs; e < grp.nst; e += NF, dst += rstep) {
                    double v[4];
                    row(e, v);
                    for (int ka = 0; ka <= nku; ++ka)
                        for (int kc = 0; kc <= nw; ++kc) {
                            double x[4] = {v[0], v[1], v[2], v[3]}, f[4];
                            if (kc) {
                                lq_ld<LS>(su + (a.m + kc - 1) * LS, t2, f);
#pragma unroll
                                for (int t = 0; t < 4; ++t) x[t] = KF_MUL(f[t], x[t]);
                            }
                            if (ka) {
                                lq_ld<LS>(su + (ka - 1) * LS, t2, f);
#pragma unroll
                                for (int t = 0; t < 4; ++t) x[t] = KF_MUL(f[t], x[t]);
                            }
                            put(dst + (long long)(ka * (nw + 1) + kc) * bstep, x);
                        }
                }
            }
            if (a.mode == 1 && a.model == KF_LINEAR && g == 0) {   // [psi_L, u]  (Ksysid.m:1062-1063)
                for (int i = fs; i < a.m; i += NF) {
                    double x[4];
                    lq_ld<LS>(su + i * LS, t2, x);
                    put(base + ((long long)a.N * (a.nw + 1) + i) * a.ld, x);
                }
            }
        }
        // (the barrier at the top of the next iteration separates this tile's readers from the next tile's writers)
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// launches the tile kernel if the program fits (levels, shared memory); returns false otherwise
template <int LS>
bool lift_tile_launch_ls(kf_ctx* ctx, KfLiftTileArgs& a, int nsides, size_t smem, cudaStream_t st, int* rc) {
    if (kf_ensure_smem(ctx, kf_lift_tile_kernel<LS>, smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    const long long ntiles = (a.M + LS - 1) / LS;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (227 * 1024) / std::max<size_t>(smem + 1024, 1)));
    const int rows = std::max(1, (ctx->sm_count * per_sm) / std::max(1, nsides * a.ngroups));
    const unsigned gx = (unsigned)std::min<long long>(ntiles, (long long)rows);
    kf_lift_tile_kernel<LS><<<dim3(gx, nsides * a.ngroups), LT_THREADS, smem, st>>>(a);
    if (cudaGetLastError() != cudaSuccess) { *rc = KF_ECUDA; ctx->err = "kf_lift_tile_kernel launch failed"; }
    ctx->launches += 1;
    return true;
}

// group tables (ops | store lists | group records) of ctx->lt_* -> device; the previous tables may still be read by kernels on
// another stream of this context, and the new ones must be visible to every stream that launches an evaluator afterwards
bool lift_groups_upload(kf_ctx* ctx, cudaStream_t st, int* rc) {
    const size_t nb_ops = ctx->lt_ops.size() * sizeof(LtOp), nb_st = ctx->lt_store.size() * sizeof(LtStore), nb_g = ctx->lt_groups.size() * sizeof(LtGroup);
    const size_t o_st = (nb_ops + 15) & ~(size_t)15, o_g = (o_st + nb_st + 15) & ~(size_t)15;
    // (this context's streams only: a device-wide synchronisation would invalidate a CUDA-graph capture running in another
    // thread's context on the same device)
    auto sync_own = [&]() -> bool {
        cudaStream_t own[] = {ctx->stream, ctx->stream2, ctx->stream_x[0], ctx->stream_x[1]};
        for (cudaStream_t q : own)
            if (q && cudaStreamSynchronize(q) != cudaSuccess) return false;
        return true;
    };
    if (!sync_own() || ctx->d_lift_groups.ensure(o_g + nb_g) != cudaSuccess) { cudaGetLastError(); return false; }
    char* b0 = ctx->d_lift_groups.as<char>();
    cudaMemcpyAsync(b0, ctx->lt_ops.data(), nb_ops, cudaMemcpyHostToDevice, st);
    if (nb_st) cudaMemcpyAsync(b0 + o_st, ctx->lt_store.data(), nb_st, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(b0 + o_g, ctx->lt_groups.data(), nb_g, cudaMemcpyHostToDevice, st);
    if (cudaStreamSynchronize(st) != cudaSuccess || !sync_own()) {
        *rc = KF_ECUDA;
        ctx->err = "lift groups upload failed";
        return false;
    }
    return true;
}
void lift_groups_bind(kf_ctx* ctx, KfLiftTileArgs& a) {
    const size_t nb_ops = ctx->lt_ops.size() * sizeof(LtOp), nb_st = ctx->lt_store.size() * sizeof(LtStore);
    const size_t o_st = (nb_ops + 15) & ~(size_t)15, o_g = (o_st + nb_st + 15) & ~(size_t)15;
    char* base = ctx->d_lift_groups.as<char>();
    a.gops = reinterpret_cast<const LtOp*>(base);
    a.gstore = reinterpret_cast<const LtStore*>(base + o_st);
    a.groups = reinterpret_cast<const LtGroup*>(base + o_g);
    a.ngroups = (int)ctx->lt_groups.size();
    a.max_slots = ctx->lt_max[0]; a.max_ops = ctx->lt_max[1]; a.max_nst = ctx->lt_max[2];
}

// Builds (or reuses) the dependency-closed feature groups of the narrow tile kernel for a slot budget and makes them resident on
// the device.  false: the program cannot be grouped (dependency chain too deep) or the upload failed (*rc set).
bool lift_groups_prepare(kf_ctx* ctx, KfLiftTileArgs& a, int slots_target, bool single, cudaStream_t st, int* rc) {
    const KfProgram& p = ctx->prog;
    const unsigned long long key[4] = {ctx->prog_gen, 0ull, (unsigned long long)slots_target, single ? 1ull : 0ull};
    const bool cached = std::equal(key, key + 4, ctx->lt_key) && !ctx->lt_groups.empty();
    if (!cached) {
        ctx->lt_key[0] = ~0ull;
        if (!kf_build_lift_groups(p, slots_target, single, ctx->lt_ops, ctx->lt_store, ctx->lt_groups)) return false;
        ctx->lt_max[0] = ctx->lt_max[1] = ctx->lt_max[2] = 0;
        for (const LtGroup& g : ctx->lt_groups) {
            ctx->lt_max[0] = std::max(ctx->lt_max[0], g.nslots);
            ctx->lt_max[1] = std::max(ctx->lt_max[1], g.nops);
            ctx->lt_max[2] = std::max(ctx->lt_max[2], g.nst);
        }
        if (!lift_groups_upload(ctx, st, rc)) return false;
        std::copy(key, key + 4, ctx->lt_key);
    }
    lift_groups_bind(ctx, a);
    return true;
}

template <int LS, int MINB>
bool lift_stream_launch_ls(kf_ctx* ctx, KfLiftTileArgs& a, int nsides, size_t smem, cudaStream_t st, int* rc) {
    if (kf_ensure_smem(ctx, kf_lift_stream_kernel<LS, MINB>, smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    const long long ntiles = (a.M + LS - 1) / LS;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(MINB, (227 * 1024) / (smem + 1024)));
    const int rows = std::max(1, (ctx->sm_count * per_sm) / std::max(1, nsides * a.ngroups));
    const unsigned gx = (unsigned)std::min<long long>(ntiles, (long long)rows);
    kf_lift_stream_kernel<LS, MINB><<<dim3(gx, nsides * a.ngroups), LT_THREADS, smem, st>>>(a);
    if (cudaGetLastError() != cudaSuccess) { *rc = KF_ECUDA; ctx->err = "kf_lift_stream_kernel launch failed"; }
    ctx->launches += 1;
    return true;
}

// streaming kernel: needs 16-byte aligned vector stores (even ld, aligned base), no dim_red, enough snapshots to fill tiles
bool lift_stream_launch(kf_ctx* ctx, KfLiftTileArgs& a, int nsides, cudaStream_t st, int* rc) {
    if (!ctx->opt_lift_wide || a.n_pcs > 0 || (a.ld & 1) || (reinterpret_cast<unsigned long long>(a.out) & 15ull) || a.M < 256) return false;
    if (a.mode == 0 && nsides > 1 && (((long long)a.side_off * a.ld) & 1)) return false;
    const KfProgram& p = ctx->prog;
    if (p.ngauss > 65535) return false;
    // tile width: 256-byte runs reach the full write rate only on 128-byte-aligned rows; otherwise 512-byte runs — unless the
    // wider tile would split the dictionary into more groups (parents evaluated once per group)
    const bool aligned = (a.ld % 16 == 0) && (reinterpret_cast<unsigned long long>(a.out) % 128 == 0);
    const int n = p.n_full();
    const int nin = a.nv + a.m + a.nw;
    const double budget = std::max(16.0, ctx->opt_lift_smem_kb) * 1024.0;      // per CTA: 2 CTAs per SM
    int ls = ctx->opt_lift_ls, max_rows = 0, slots_target = 0;
    const bool ls_auto = ls != 32 && ls != 64;
    if (ls_auto) ls = aligned ? 32 : 64;
    auto plan = [&](int w) {
        const long long ntiles = (a.M + w - 1) / w;
        // few tiles (a panel of the fit path, a short series): split the rows so that every SM has work
        max_rows = 0;
        if (ntiles * nsides < 2LL * ctx->sm_count) {
            const int parts = (int)((2LL * ctx->sm_count + ntiles * nsides - 1) / (ntiles * nsides));
            max_rows = std::max(32, (n + parts - 1) / parts);
        }
        slots_target = std::max(p.nv + 8, (int)((budget - 8.0 * std::min(n, max_rows > 0 ? max_rows : n)) / (w * sizeof(double))) - a.m - a.nw - nin);
    };
    plan(ls);
    if (ls_auto && ls == 64) {
        // cached decision per (program, budget): does the 64-wide plan need more groups than the 32-wide one?
        const unsigned long long pk[3] = {ctx->prog_gen, (unsigned long long)budget, (unsigned long long)(a.m + a.nw)};
        if (!std::equal(pk, pk + 3, ctx->lt_plan_key)) {
            std::vector<LtOp> ops_tmp;
            std::vector<LtGroup> g64, g32;
            const bool ok64 = kf_build_lift_rowgroups(p, slots_target, 0, ops_tmp, g64);
            const int st64 = slots_target;
            plan(32);
            const bool ok32 = kf_build_lift_rowgroups(p, slots_target, 0, ops_tmp, g32);
            ctx->lt_plan_narrow = ok32 && (!ok64 || g32.size() < g64.size());
            std::copy(pk, pk + 3, ctx->lt_plan_key);
            (void)st64;
        }
        ls = ctx->lt_plan_narrow ? 32 : 64;
        plan(ls);
    }
    auto bytes = [&](int slots, int ops, int nst, int ncent) {
        return ((size_t)slots + (size_t)a.m + (size_t)a.nw + (size_t)nin) * ls * sizeof(double) + (size_t)ncent * sizeof(double) +
               (size_t)ops * sizeof(LtOp) + (size_t)nst * sizeof(LtRop) + (size_t)nin * sizeof(double*) + 64;
    };
    {
        const unsigned long long key[4] = {ctx->prog_gen, 1ull, (unsigned long long)slots_target, (unsigned long long)max_rows};
        const bool cached = std::equal(key, key + 4, ctx->lt_key) && !ctx->lt_groups.empty();
        if (!cached) {
            ctx->lt_key[0] = ~0ull;
            if (!kf_build_lift_rowgroups(p, slots_target, max_rows, ctx->lt_ops, ctx->lt_groups)) return false;
            ctx->lt_max[0] = ctx->lt_max[1] = ctx->lt_max[2] = 0;
            for (const LtGroup& g : ctx->lt_groups) {
                if (g.nslots > 65535) return false;
                ctx->lt_max[0] = std::max(ctx->lt_max[0], g.nslots);
                ctx->lt_max[1] = std::max(ctx->lt_max[1], g.nops);
                ctx->lt_max[2] = std::max(ctx->lt_max[2], g.nst);
            }
            // the row ops packed to 8 bytes (indexed by output row) | the constants of the primitive ops, in the `store` segment
            std::vector<double> consts;
            std::vector<LtRop> rops((size_t)n);
            for (const LtGroup& g : ctx->lt_groups)
                for (int e = 0; e < g.nst; ++e) {
                    const LtOp& o = ctx->lt_ops[(size_t)g.rop_off + e];
                    LtRop r{};
                    r.kind = (unsigned char)o.kind; r.a = (unsigned short)o.a; r.b = (unsigned short)o.b;
                    if (o.kind != KF_OP_MUL && o.kind != KF_OP_VAR && o.kind != KF_OP_GAUSS) {
                        size_t ci = std::find(consts.begin(), consts.end(), o.c) - consts.begin();
                        if (ci == consts.size()) consts.push_back(o.c);
                        if (ci > 65535) return false;
                        r.ci = (unsigned short)ci;
                    }
                    rops[(size_t)g.row0 + e] = r;
                }
            if (consts.empty()) consts.push_back(0.0);
            static_assert(sizeof(LtStore) == 8 && sizeof(LtRop) == 8, "the packed row ops travel in the store segment");
            ctx->lt_store.resize(rops.size() + consts.size());
            std::memcpy(ctx->lt_store.data(), rops.data(), rops.size() * sizeof(LtRop));
            std::memcpy(ctx->lt_store.data() + rops.size(), consts.data(), consts.size() * sizeof(double));
            if (!lift_groups_upload(ctx, st, rc)) return *rc != KF_OK;
            std::copy(key, key + 4, ctx->lt_key);
        }
        lift_groups_bind(ctx, a);
        a.rops = reinterpret_cast<const unsigned long long*>(a.gstore);
        a.rconst = reinterpret_cast<const double*>(a.gstore) + n;
    }
    const int ncent_all = p.ngauss * p.nv;
    a.ncent = (ncent_all > 0 && bytes(a.max_slots, a.max_ops, a.max_nst, ncent_all) <= (size_t)std::max(budget, 100.0 * 1024)) ? ncent_all : 0;
    const size_t b = bytes(a.max_slots, a.max_ops, a.max_nst, a.ncent);
    if (b > 200 * 1024) return false;                    // does not fit: the narrow kernel takes it
    if (ctx->opt_lift_minb >= 3)
        return ls == 64 ? lift_stream_launch_ls<64, 3>(ctx, a, nsides, b, st, rc) : lift_stream_launch_ls<32, 3>(ctx, a, nsides, b, st, rc);
    return ls == 64 ? lift_stream_launch_ls<64, 2>(ctx, a, nsides, b, st, rc) : lift_stream_launch_ls<32, 2>(ctx, a, nsides, b, st, rc);
}

bool lift_tile_launch(kf_ctx* ctx, KfLiftTileArgs& a, int nsides, cudaStream_t st, int* rc) {
    *rc = KF_OK;
    const int nlev = (int)ctx->level_start.size() - 1;
    if (nlev > LT_MAXLEV || a.M <= 0) return false;
    if (lift_stream_launch(ctx, a, nsides, st, rc)) return true;
    if (*rc) return true;
    const KfProgram& p = ctx->prog;
    auto bytes = [&](int ls, int slots, int ops, int nst) {
        return ((size_t)slots + (size_t)a.m + (size_t)a.nw + (a.n_pcs > 0 ? (size_t)a.N : 0)) * ls * sizeof(double) + (size_t)ops * sizeof(LtOp) +
               (size_t)nst * sizeof(LtStore) + 64;
    };
    // Narrow tiles (8 / 16 snapshots; dim_red, odd leading dimensions, few snapshots).  Option lift_ls forces the width.
    int ls = ctx->opt_lift_ls;
    if (ls != 8 && ls != 16 && ls != 32 && ls != 64) ls = 16;
    if (a.n_pcs > 0 && ls > 16) ls = 16;                 // dim_red needs every feature in one CTA
    const int slots_target = std::max(p.nv + 8, (int)((56 * 1024) / (ls * sizeof(double))));
    const bool single = a.n_pcs > 0 || bytes(ls, p.n_full(), p.n_full(), p.n_full()) <= 74 * 1024;
    if (!lift_groups_prepare(ctx, a, slots_target, single, st, rc)) return *rc != KF_OK;
    for (;;) {
        const size_t b = bytes(ls, a.max_slots, a.max_ops, a.max_nst);
        if (b <= 200 * 1024) {
            if (ls == 64) return lift_tile_launch_ls<64>(ctx, a, nsides, b, st, rc);
            if (ls == 32) return lift_tile_launch_ls<32>(ctx, a, nsides, b, st, rc);
            if (ls == 16) return lift_tile_launch_ls<16>(ctx, a, nsides, b, st, rc);
            return lift_tile_launch_ls<8>(ctx, a, nsides, b, st, rc);
        }
        if (ls == 8) break;
        ls /= 2;                                         // a single group (dim_red, or one huge closure) that does not fit: narrower tile
    }
    return false;
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
    static bool carve = false;
    if (!carve) {      // run beside the INT8 contraction of the other chunk pipeline (see kf_oz_prepare)
        cudaFuncSetAttribute(kf_lift_level_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(kf_lift_econ_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaGetLastError();
        carve = true;
    }
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

// the points [c0, c0 + cnt) of V (total x nv, ld = total) -> out (N x cnt, ld = ldo); no dim_red
int kf_launch_lift_points_chunk(kf_ctx* ctx, const KfOp* ops, const double* centres, int nv, int n_full, const double* V, long long total,
                                long long c0, long long cnt, double* out, long long ldo, cudaStream_t st) {
    if (cnt <= 0) return KF_OK;
    if (ctx->opt_lift_tile) {
        KfLiftTileArgs t{};
        t.ops = ops; t.centres = centres; t.pcs = nullptr; t.order = ctx->d_order.as<int>();
        t.nv = nv; t.n_full = n_full; t.n_pcs = 0; t.N = n_full;
        t.nzeta = nv; t.m = 0; t.model = KF_NONLINEAR; t.mode = 0; t.P = t.N;
        t.alpha = V + c0; t.beta = V + c0; t.u = V + c0; t.M = cnt; t.ldin = total; t.out = out; t.ld = ldo;
        int rc = KF_OK;
        if (lift_tile_launch(ctx, t, 1, st, &rc)) return rc;
    }
    KfLiftArgs a{};
    a.ops = ops; a.centres = centres; a.pcs = nullptr; a.order = ctx->d_order.as<int>();
    a.nv = nv; a.n_full = n_full; a.n_pcs = 0; a.N = n_full;
    a.nzeta = nv; a.m = 0; a.model = KF_NONLINEAR;
    a.alpha = V; a.beta = V; a.u = V;
    if (c0 + cnt != total) {       // the level kernel uses M both as the input leading dimension and as the end of the data
        ctx->err = "chunked point lift: this dictionary needs the level-by-level kernel, which only takes the last chunk";
        return KF_EUNSUPPORTED;
    }
    a.M = total; a.start = c0; a.Mc = (int)cnt;
    a.panel = out; a.ld = ldo; a.full = nullptr;
    a.x_off = 0; a.y_off = 0; a.w_off = 0; a.nW = 0;
    a.nsides = 1; a.extras = 0;
    return kf_launch_lift(ctx, a, st);
}

int kf_launch_lift_points(kf_ctx* ctx, const KfOp* ops, const double* centres, const double* pcs, int nv, int n_full,
                          int n_pcs, const double* V, long long rows, double* full, double* out, long long ldo,
                          cudaStream_t st) {
    if (rows <= 0) return KF_OK;
    if (ctx->opt_lift_tile) {
        KfLiftTileArgs t{};
        t.ops = ops; t.centres = centres; t.pcs = pcs; t.order = ctx->d_order.as<int>();
        t.nv = nv; t.n_full = n_full; t.n_pcs = n_pcs; t.N = n_pcs ? nv + n_pcs + 1 : n_full;
        t.nzeta = nv; t.m = 0; t.model = KF_NONLINEAR; t.mode = 0; t.P = t.N;
        t.alpha = V; t.beta = V; t.u = V; t.M = rows; t.ldin = rows; t.out = out; t.ld = ldo;
        int rc = KF_OK;
        if (lift_tile_launch(ctx, t, 1, st, &rc)) return rc;
    }
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

// ---------------------------------------------------------------------------------------------------------------
// Panel lift through the shared-memory tile evaluator: psi(alpha) -> panel rows [x_off, x_off + N), psi(beta) -> rows
// [y_off, y_off + N) for the chunk [start, start + Mc) (feature-major, ld = Mc, L2-resident), plus the u rows (linear)
// and the Kronecker weight rows u_a u_b (bilinear).  ~3x the rate of the level-by-level kernel, which re-reads product
// factors from the panel and pays one launch per dependency level.
__global__ void kf_panel_extras_kernel(const KfLiftArgs a, long long count) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.Mc) return;
    const bool valid = s < count;
    const long long gs = a.start + s;
    if (a.model == KF_LINEAR) {
        double* sec = a.panel + (long long)a.x_off * a.ld + s;
        for (int i = 0; i < a.m; ++i) sec[(long long)(a.N + i) * a.ld] = valid ? a.u[(long long)i * a.M + gs] : 0.0;
    } else if (a.model == KF_BILINEAR && a.nW > 0) {
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

// returns false if the tile evaluator cannot take the program (then the caller uses the level kernel)
bool kf_launch_lift_panel_tile(kf_ctx* ctx, const KfLiftArgs& a, cudaStream_t st, int* rc, long long limit) {
    *rc = KF_OK;
    long long count = std::min<long long>(a.Mc, a.M - a.start);
    if (limit >= 0) count = std::min(count, limit);      // snapshots of the chunk that belong to this call
    if (count <= 0 || !ctx->opt_lift_tile) return false;
    KfLiftTileArgs t{};
    t.ops = a.ops; t.centres = a.centres; t.pcs = a.pcs; t.order = a.order;
    t.nv = a.nv; t.n_full = a.n_full; t.n_pcs = a.n_pcs; t.N = a.N;
    t.nzeta = a.nzeta; t.m = 0; t.nw = 0; t.model = a.model; t.mode = 0; t.P = a.N;   // variables beyond nzeta (nonlinear) come from u
    t.alpha = a.alpha + a.start; t.beta = a.beta + a.start; t.u = a.u ? a.u + a.start : a.u; t.M = count; t.ldin = a.M;
    t.out = a.panel + (long long)a.x_off * a.ld; t.ld = a.ld; t.side_off = a.y_off - a.x_off;
    if (count < a.Mc) {                              // tail of the last chunk: the tile kernel stores valid snapshots only
        const size_t bytes = (size_t)(a.y_off + a.N - a.x_off) * a.ld * sizeof(double);
        if (cudaMemsetAsync(a.panel + (long long)a.x_off * a.ld, 0, bytes, st) != cudaSuccess) { *rc = KF_ECUDA; return true; }
    }
    if (!lift_tile_launch(ctx, t, a.nsides > 0 ? a.nsides : 2, st, rc)) return false;
    if (*rc) return true;
    if (a.extras && (a.model == KF_LINEAR || (a.model == KF_BILINEAR && a.nW > 0))) {
        kf_panel_extras_kernel<<<(a.Mc + 255) / 256, 256, 0, st>>>(a, count);
        if (cudaGetLastError() != cudaSuccess) { *rc = KF_ECUDA; ctx->err = "kf_panel_extras_kernel launch failed"; }
        ctx->launches += 1;
    }
    return true;
}

// Materialised regressors of the snapshots [a0.start, a0.start + count) (count = min(a0.Mc, a0.M - a0.start); a0.Mc <= 0: all):
// psi(x) -> columns [0,N), psi(y) -> columns [P, P+N) of AB (ld = ldp >= count), plus the u / u (x) psi columns.
int kf_launch_regressors(kf_ctx* ctx, const KfLiftArgs& a0, double* AB, double* /*unused*/, long long ldp,
                         cudaStream_t st) {
    KfLiftArgs a = a0;
    if (a.Mc <= 0) { a.start = 0; a.Mc = (int)a.M; }
    const long long count = std::min<long long>(a.Mc, a.M - a.start);
    if (count <= 0) return KF_OK;
    const int P = kf_regressor_width(a.model, a.N, a.m, a.nw);
    if (ctx->opt_lift_tile) {
        KfLiftTileArgs t{};
        t.ops = a.ops; t.centres = a.centres; t.pcs = a.pcs; t.order = a.order;
        t.nv = a.nv; t.n_full = a.n_full; t.n_pcs = a.n_pcs; t.N = a.N;
        t.nzeta = a.nzeta; t.m = a.m; t.model = a.model; t.mode = 1; t.P = P;
        t.alpha = a.alpha + a.start; t.beta = a.beta + a.start; t.u = a.u ? a.u + a.start : a.u; t.M = count; t.ldin = a.M;
        t.w = (a.nw > 0 && a.w) ? a.w + a.start : nullptr; t.nw = a.nw;
        t.out = AB; t.ld = ldp;
        int rc = KF_OK;
        if (lift_tile_launch(ctx, t, 2, st, &rc)) return rc;
    }
    a.panel = AB;
    a.ld = ldp;
    a.Mc = (int)count;
    a.x_off = 0;
    a.y_off = P;
    a.nW = 0;
    a.w_off = 0;
    a.nsides = 2;
    a.extras = 0;
    KF_TRY(kf_launch_lift(ctx, a, st));
    if (a.model != KF_NONLINEAR || a.nw > 0) {
        dim3 grid((unsigned)((count + 127) / 128), (a.model == KF_BILINEAR || a.nw > 0) ? (a.N + 63) / 64 : 1);
        kf_regressor_post_kernel<<<grid, 128, 0, st>>>(a.model, a.N, P, a.m, a.nw, a.u ? a.u + a.start : a.u, a.nw > 0 ? a.w + a.start : nullptr,
                                                       count, a.M, AB, ldp);
        KF_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    return KF_OK;
}
