// Internal declarations shared by the .cu files of libkoopfit.so (not part of the ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <utility>
#include <string>
#include <vector>

#include "../../include/koopfit.h"
#include "program.h"

// ------------------------------------------------------------------ tiling constants
constexpr int KF_BM = 128;          // output tile rows  (A-operand rows per CTA)
constexpr int KF_BN = 128;          // output tile cols  (B-operand rows per CTA)
constexpr int KF_BK = 16;           // contraction elements per pipeline stage
constexpr int KF_STAGES = 3;
// CTA tile of the production DMMA kernel: 128 x 64, 4 warps (2 x 2, warp tile 64 x 32), 2 CTAs per SM —
// the best of the configurations measured in profiles/r01_gemm_variants_v2.txt
constexpr int KF_CTA_M = 128;
constexpr int KF_CTA_N = 64;
constexpr int KF_GEMM_THREADS = 128;
constexpr int KF_TILE_ELEMS = KF_BM * KF_BN;

inline long long kf_roundup(long long x, long long m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------ GEMM task
// One CTA computes  out[m*ldm + n*ldn] (+)= alpha * sum_{k in [k0,k1)} w[k] * A[m*lda + k] * B[n*ldb + k]
// for a BM x BN tile.  Both operands are "k-contiguous": row r of the operand holds the
// contraction index contiguously (a lifted-panel row = one observable over the snapshots
// of the chunk; a column of a column-major matrix).  k0,k1 multiples of KF_BK; rows beyond
// a_rows/b_rows are clamped for loading and masked on store.
struct KfGemmTask {
    const double* A;
    const double* B;
    const double* W;      // per-k weights (NULL = none): bilinear u_a*u_b factor
    double* out;
    long long lda, ldb;   // operand row strides (doubles), multiples of 2
    long long ldm, ldn;   // output strides
    int k0, k1;
    int a_rows, b_rows;   // valid rows of this tile (<= BM / BN)
    double alpha;
    int accumulate;       // 1: out += ; 0: out =
    int pad;
};

// Task of the tensor-map TMA Gram kernel: operands are addressed by panel ROW (the panel itself is described by
// a CUtensorMap), one CTA = one 128 x 64 sub-tile of a 128 x 128 accumulator tile.
struct KfTmaTask {
    int a_row, b_row, w_row;   // first panel row of the A / B operand; w_row unused (kept for debugging)
    int k0, k1;                // snapshot range inside the panel, multiples of 16
    int pad;
    double* out;               // sub-tile base inside the accumulator tile (row stride 128)
    const double* W;           // weight row (global pointer) or nullptr = unweighted
};

// ------------------------------------------------------------------ device buffer
struct KfBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ------------------------------------------------------------------ accumulation layout
struct KfTile {
    int kind;   // 0 = G tile, 1 = C tile
    int q;      // weight row (bilinear pair index; 0 = unweighted)
    int a, b;   // bilinear block coordinates (u_a, u_b), a <= b; 0,0 otherwise
    int tm, tn; // tile coordinates inside the block
};

struct KfLayout {
    int model = 0, m = 0, nzeta = 0, nv = 0, nw = 0;
    bool dense = false;       // G, C accumulated densely (rf.d_dense): `loaded` models, and the INT8 engine
    bool oz = false;          // INT8 tensor-core engine (ozaki.cu): accumulators are ROW-major [G | C] per chunk pipeline
    int n_full = 0, N = 0, P = 0;
    int Pc = 0;               // columns of C / K actually computed (pc_cols fast mode), <= P
    int Rx = 0, Rxp = 0;      // rows of the X section: N (+m for linear), padded to BM
    int Ny = 0, Nyp = 0;      // rows of the Y section (N), padded
    int nW = 0;               // weight rows (bilinear): (m+1)(m+2)/2
    int x_off = 0, y_off = 0, w_off = 0, rows = 0;   // panel row offsets / total rows
    int Mc = 0;               // snapshots per panel (multiple of KF_BK)
    int npipes = 2;           // chunk pipelines in flight
    int nsplit = 1;           // split-K slabs
    std::vector<KfTile> tiles;
    int Pp = 0;               // P padded to BM: leading dimension of G, C, K work matrices
    long long slab = 0;       // doubles per accumulator slab: tiles * 128 * 128 + KF_ACC_TRAILER (the trailer of slab 0 carries
                              // the snapshot count through the all-reduce)
    bool valid = false;
};
constexpr int KF_ACC_TRAILER = 16;
constexpr int KF_MAX_PIPES = 4;     // chunk pipelines of the INT8 engine (the DMMA path uses two)

// state of the Gram-route refinement (multi-level pivoted Cholesky-QR; api.cu: gram_ls_step / refine_pass)
struct KfRefine {
    bool pending = false;     // the solve asked for another data pass; kf_accumulate_dev(reset = 0) then runs refine_pass
    int level = 0;            // refinement passes done
    int forced = 0;           // features accepted (and orthonormalised) so far
    int Mc = 0;               // snapshots per chunk of a refinement pass
    double r11 = 0;           // |R_11| of the first factorisation (MATLAB's rank tolerance refers to it)
    double min_piv = 0;       // smallest accepted |R_jj| so far (true scale)
    double cond_est = 0;
    long long M_total = 0;    // snapshots over all ranks
    std::vector<int> orig;    // original regressor column of feature j
    KfBuf d_S, d_St, d_Sp;    // basis S (row = new feature, column = original; column-major), its transpose, Pi S
    KfBuf d_G2C2;             // [G2 | C2] of the new features: what the ranks all-reduce in a refinement pass
    KfBuf d_RP, d_Z;          // chunk panels: materialised [Px | Py] rows and the transformed features
    KfBuf d_dense;            // dense accumulator [G | C | trailer] (2 Pp^2 + KF_ACC_TRAILER doubles): `loaded` models, INT8 engine
    KfBuf d_dense2;           // INT8 engine: the accumulator set of the second chunk pipeline
    KfBuf d_dense_x[2];       // ... of pipelines 2, 3
};

// the blocked pivoted Cholesky of kf_solve_gram_ls as an instantiated CUDA graph, valid for one set of buffers / sizes
struct KfPcholGraph {
    cudaGraphExec_t exec = nullptr;
    const void *W = nullptr, *perm = nullptr, *dcur = nullptr, *state = nullptr;
    int P = 0, Pp = 0;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    double flops = 0;
};

struct KfOzState;   // ozaki.cu

struct kf_ctx {
    int device = 0;
    KfOzState* oz = nullptr;
    int opt_oz_sym = 1;       // INT8 engine: exploit the Kronecker block symmetry of a bilinear regressor
    int opt_gram_engine = 0;  // 0 auto, 1 FP64 DMMA, 2 INT8 tensor cores (Ozaki scheme II, FP64-exact)
    double i8_ops = 0;        // INT8 tensor-core operations issued
    double i8_ops_per_launch = 0;
    int last_engine = 1;      // engine of the last accumulate: 1 DMMA, 2 INT8
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaStream_t stream_x[2] = {nullptr, nullptr};     // pipelines 2, 3 of the INT8 engine
    cudaEvent_t ev_join[KF_MAX_PIPES] = {};
    int opt_oz_pipes = 2;     // INT8 engine: chunk pipelines in flight (2..4; measured: 3 and 4 give nothing over 2 — the element-wise
                              // kernels do not run beside the persistent contraction, profiles/r02_int8_engine_summary.md)
    cudaEvent_t ev[9] = {};
    cudaEvent_t ev_fork = nullptr;
    cudaStream_t copy_stream = nullptr;      // host -> device copies of kf_fit, overlapped with the lift + Gram
    std::vector<cudaEvent_t> copy_ev;
    // NCCL communicator of a multi-GPU fit (comm.cu): one rank per context; nullptr = single GPU
    void* comm = nullptr;
    int nranks = 1, rank = 0;
    int comm_owned = 0;
    std::string err;
    int sm_count = 148;

    KfProgram prog;
    KfLayout lay;

    std::vector<int> level_start;   // offsets into the level-sorted feature order
    KfBuf d_order;
    KfBuf d_ops, d_centres, d_pcs, d_panel[KF_MAX_PIPES], d_full, d_tasks[2], d_tma_tasks[2], d_accum, d_tilemeta;
    CUtensorMap tmap[2];            // tensor maps of the two panels (SWIZZLE_128B, box 16 x 64)
    KfBuf d_G, d_C, d_K, d_W, d_in, d_misc, d_qr, d_tmp, d_K2, d_K3, d_Kt;
    int opt_lift_panel_fit = 1;          // fit path: lift the panel with the shared-memory tile evaluator instead of the level kernel
    int opt_qr_blocked = 1;              // QRCP route: blocked Householder QR of the tall matrix first (BLAS-3), pivoted QR of its R factor after
    int opt_qr_ksplit = 64;              // ... two-level summation of the products over the rows: number of k chunks (1: off)
    int opt_qr_nb = 64;                  // ... panel width (32 | 64 | 128)
    int opt_lift_wide = 1;               // materialising lift: wide tiles (32 / 64 snapshots: long DRAM runs, table-free stores)
    int opt_lift_minb = 2;               // ... resident CTAs per SM the kernel is compiled for (2: 128 registers, 3: 80)
    double opt_lift_smem_kb = 110;       // ... shared memory per CTA (decides the size of the feature groups and the CTAs per SM)
    int last_pca_sweeps = 0;
    KfBuf d_pca;                         // PCA: lifted chunk, centred Gram (+ split-K slabs), Jacobi work matrices
    KfBuf d_deal;                        // lasso sweep split by columns: C with its columns dealt round-robin to the ranks / scratch
    KfBuf d_bqr;                         // blocked QR: V panel, S, T, W, W2, tau
    KfBuf d_lift_groups;                 // feature groups of the materialising lift (ops | store lists | group records)
    KfBuf d_series;                      // raw merged series t | y | u and the scale factors (kf_fit_series)
    KfBuf d_as_mat, d_as_aux, d_as_ws;   // active-set QP solver: pattern/solution matrices, index lists, Cholesky factors

    // options
    int opt_chunk = 0;        // 0 = auto
    int opt_splitk = 0;       // 0 = auto
    int opt_overlap = 1;      // even / odd chunks as two independent lift -> Gram pipelines on two streams
    double opt_panel_mb = 64; // target size of one L2-resident panel (24/48/64 MB measured: 31.7/32.5/32.8 TF)
    double opt_qr_max_gb = 16; // KF_LS_AUTO takes the QRCP route when [Px|Py] is at most this large
    int opt_profile = 0;      // sample Gram-kernel durations with CUDA events (adds syncs)
    int opt_qp_method = 0;    // L1-ball QP: 0 auto = 2 (exact active set), 1 coordinate descent, 2 active set
    int opt_as_level = -1;    // active-set solver: level-synchronous factorisation (-1 off (default: measured no gain), 0 auto when a rank has few columns, 1 always)
    int opt_as_skip = 1;      // active-set solver: columns whose support did not change keep their a_j, b_j (no refactorisation)
    int opt_as_diag = 0;      // active-set solver: print how much factor reuse an age-ordered support would allow (diagnostic)
    double opt_as_frac = 0.05; // active-set solver: bound on the pattern change per step, as a fraction of the support size
    double opt_as_ws_gb = 8;  // active-set solver: bound on the per-column Cholesky workspace (columns run in chunks that fit)
    int opt_lift_ls = 0;      // materialising lift: snapshots per tile (8 | 16 | 32 | 64; 0 = auto)
    int opt_lift_tile = 1;    // materialising lift: whole program per 16-snapshot tile in shared memory (0: level-by-level kernel)
    int opt_tma = 1;          // Gram kernel operand path: 1 = tensor-map TMA + mbarrier ring, 0 = per-thread cp.async
    int opt_qp_split = 1;     // multi-GPU context: split the active-set lasso sweep by columns of K over the ranks
    int opt_refine = 1;       // Gram-route refinement: 0 off, 1 adaptive, 2 always one extra pass
    double opt_refine_kappa = 1e3;      // refine when the pivot ratio of the first factorisation exceeds this
    double opt_refine_level_tol = 1e-5; // dynamic range of |R_jj| one level resolves (its square must stay well above eps)
    int opt_refine_max = 4;   // level limit
    KfRefine rf;
    double accum_M_d = 0;     // accum_M as a double (source of the accumulator trailer upload)

    // host staging of the lift feature groups (kept alive across the asynchronous upload)
    std::vector<LtOp> lt_ops;
    std::vector<LtStore> lt_store;
    std::vector<LtGroup> lt_groups;
    // the uploaded group tables are reused while (program, tile width, group size) stay the same: the panel pipeline launches
    // the tile evaluator once per chunk and must not pay a host synchronisation each time
    unsigned long long prog_gen = 0;     // bumped by prepare_program when the dictionary changed
    unsigned long long prog_hash = 0;    // FNV-1a of the compiled dictionary
    unsigned long long lt_key[4] = {~0ull, 0, 0, 0};
    int lt_max[3] = {0, 0, 0};           // max_slots, max_ops, max_nst of the cached groups
    unsigned long long lt_plan_key[3] = {~0ull, 0, 0};   // streaming lift: cached tile-width decision for unaligned rows
    bool lt_plan_narrow = false;

    // column partition of the active-set QP solver across ranks (kf_set_qp_partition)
    int qp_lo = 0, qp_hi = 0;
    kf_allreduce_fn qp_allreduce = nullptr;
    void* qp_user = nullptr;


    KfPcholGraph pchol_graph;
    int opt_graphs = 1;       // replay the launch-bound pivoted-Cholesky loop as a CUDA graph

    // counters
    double dmma_flops = 0;
    long long launches = 0;
    float last_lift_gram_ms = 0, last_gram_kernel_ms = 0, last_solve_ms = 0, last_refine_ms = 0;
    float gram_ms_sampled = 0;
    int gram_samples = 0;
    long long last_gram_launches = 0;
    int last_gram_sampled = 0;
    float last_allreduce_ms = 0;
    long long accum_M = 0;    // snapshots accumulated since reset
};

#define KF_CUDA(ctx, call)                                                                     \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                         ":" + std::to_string(__LINE__) + ")";                                 \
            return KF_ECUDA;                                                                   \
        }                                                                                      \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize of `func` raised to at least `bytes` on the context's device
std::mutex& kf_smem_table_mutex();
std::map<std::pair<int, const void*>, size_t>& kf_smem_table();
template <class F>
inline cudaError_t kf_ensure_smem(kf_ctx* ctx, F* func, size_t bytes) {
    // The attribute is per DEVICE and function, not per context: several contexts (Fitter objects, the per-device contexts of
    // kf_create_multi) share it, and a context that set a smaller value after another one's larger request would make that
    // one's next launch fail with "invalid argument".  Hence a process-wide table that only ever grows (api.cu).
    std::lock_guard<std::mutex> lock(kf_smem_table_mutex());
    size_t& have = kf_smem_table()[std::make_pair(ctx->device, reinterpret_cast<const void*>(func))];
    if (have == 0) have = 48 * 1024;
    if (bytes <= have) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

#define KF_TRY(expr)            \
    do {                        \
        int rc__ = (expr);      \
        if (rc__) return rc__;  \
    } while (0)

// ------------------------------------------------------------------ kernels (host launchers)
// gemm.cu
int kf_launch_gemm_tasks(kf_ctx* ctx, const KfGemmTask* d_tasks, int ntasks, bool weighted, cudaStream_t st);
int kf_launch_gram_tma(kf_ctx* ctx, const KfTmaTask* d_tasks, int ntasks, bool weighted, const CUtensorMap& tmap, cudaStream_t st);
// grid form: out(M x N tiles) over column-major / k-contiguous operands
struct KfGemmGrid {
    const double* A; const double* B; double* out;
    long long lda, ldb, ldm, ldn;
    int m, n;          // output extent (rows of A / rows of B)
    int k0, k1;
    double alpha; int accumulate;
    int lower_only;    // 1: skip tiles strictly above the diagonal (tn > tm)
    int ksplit;        // > 1 (kf_launch_gemm_grid, accumulate = 0): the k range is cut into ksplit chunks, chunk z writes its partial
    long long slab;    //     product to out + z * slab; the caller sums the slabs (kf_reduce_slabs) — a two-level summation
};
int kf_launch_gemm_grid(kf_ctx* ctx, const KfGemmGrid& g, cudaStream_t st);
// number of non-empty k chunks (multiples of KF_BK) when a range of K is cut into about `want` pieces
inline int kf_gemm_ksplit(int K, int want) {
    if (want <= 1 || K <= KF_BK) return 1;
    const int kc = (int)kf_roundup((K + want - 1) / want, KF_BK);
    return (K + kc - 1) / kc;
}
int kf_launch_gemm_bkmajor(kf_ctx* ctx, const KfGemmGrid& g, cudaStream_t st);   // B stored k-major: B[k * ldb + n]

// lift.cu
struct KfLiftArgs {
    const KfOp* ops; const double* centres; const double* pcs;
    const int* order;              // features sorted by dependency level
    int nsides, extras;            // 1|2 sides; extras: also write u rows / bilinear weight rows
    int nv, n_full, n_pcs, N;
    int nzeta, m, model;
    const double* alpha; const double* beta; const double* u;   // device, column-major, ld = M
    const double* w; int nw;                                    // loads (M x nw, ld = M) of a `loaded` model, or nullptr / 0
    long long M, start;   // chunk = snapshots [start, start+Mc)
    int Mc;
    double* panel; long long ld;   // panel row stride (= Mc)
    double* full;                  // scratch for dim_red: 2 * n_full rows x ld
    int x_off, y_off, w_off, nW;
};
void kf_program_levels(const KfProgram& p, std::vector<int>& order, std::vector<int>& level_start);
int kf_launch_lift(kf_ctx* ctx, const KfLiftArgs& a, cudaStream_t st);
// the panel lift through the shared-memory tile evaluator; false: not applicable (use kf_launch_lift)
bool kf_launch_lift_panel_tile(kf_ctx* ctx, const KfLiftArgs& a, cudaStream_t st, int* rc, long long limit = -1);
int kf_launch_lift_points_chunk(kf_ctx* ctx, const KfOp* ops, const double* centres, int nv, int n_full, const double* V, long long total,
                                long long c0, long long cnt, double* out, long long ldo, cudaStream_t st);
// pca.cu: principal components of the lifted points (device eigensolver)
int kf_pca_points(kf_ctx* ctx, long long rows, const double* d_V, double* mu, double* latent, double* coeff);
// materialised lift of arbitrary points: V (rows x nv, ld=rows) -> Psi (rows x N, ld = ldo)
int kf_launch_lift_points(kf_ctx* ctx, const KfOp* ops, const double* centres, const double* pcs, int nv, int n_full,
                          int n_pcs, const double* V, long long rows, double* full, double* out, long long ldo,
                          cudaStream_t st);
// materialised regressors Px, Py (M x P, ld = ldp) from device snapshot pairs
int kf_launch_regressors(kf_ctx* ctx, const KfLiftArgs& a, double* Px, double* Py, long long ldp, cudaStream_t st);

// solve.cu
int kf_reduce_slabs(kf_ctx* ctx, double* accum, long long slab_elems, int nsplit, cudaStream_t st);
int kf_assemble(kf_ctx* ctx, const double* accum, const KfTile* d_meta, int ntiles, const KfLayout& lay,
                double* G, double* C, cudaStream_t st);
// pivoted Cholesky basic solution of G K = C  (G, C, K: Pp x Pp column-major, ld = Pp)
int kf_solve_gram_ls(kf_ctx* ctx, int P, int Pp, int ncols, double* G_work, const double* C, double* K, double tol,
                     int* d_perm, int* rank_out, double* min_piv, double* max_piv, cudaStream_t st);
int kf_pchol_factor(kf_ctx* ctx, int P, int Pp, double* W, double tol, double tolabs, int forced, int* d_perm, int* rank_out,
                    double* min_piv, double* max_piv, double* pivots_host, cudaStream_t st, double* rejected_out);
int kf_rf_identity(kf_ctx* ctx, double* S, int P, int Pp, cudaStream_t st);
int kf_rf_symmetrize(kf_ctx* ctx, double* G, int Pp, cudaStream_t st);
int kf_rf_transpose(kf_ctx* ctx, const double* A, double* At, int n, cudaStream_t st);
int kf_rf_gather_rows(kf_ctx* ctx, const double* S, const int* d_perm, int P, int Pp, double* Sp, cudaStream_t st);
int kf_rf_update_basis(kf_ctx* ctx, int P, int Pp, double* W, const int* d_perm, int r, double* S, double* Sp_out, cudaStream_t st);
int kf_pchol_solve(kf_ctx* ctx, int P, int Pp, int ncols, double* W, const double* C, double* K, const int* d_perm, int r, int scatter,
                   cudaStream_t st);
// Householder QRCP basic solution of A X = B;  AB = [A | B] (M x (P+Pc), ld = ldab) is overwritten.  ldab = kf_qr_ld(M) with the
// rows [M, ldab) ZERO enables the blocked route (stage 1: blocked Householder QR on the DMMA GEMMs; stage 2: QRCP of R).
inline long long kf_qr_ld(long long M) { return kf_roundup(M, 64); }
int kf_solve_qr_ls(kf_ctx* ctx, long long M, int P, int Pc, double* AB, long long ldab, double* X, long long ldx,
                   int* d_perm, int* rank_out, double* min_piv, double* max_piv, cudaStream_t st);

// ozaki.cu: FP64-exact Gram / cross products on the INT8 tensor cores
bool kf_oz_supported(const KfLayout& L);
int kf_oz_prepare(kf_ctx* ctx, KfLayout& L);
int kf_oz_chunk(kf_ctx* ctx, const KfLayout& L, int b, const double* panel, double* accG, double* accC, cudaStream_t st,
                cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr);
int kf_oz_finish(kf_ctx* ctx, const KfLayout& L, double* acc0, double* acc1, cudaStream_t st);
int kf_oz_to_gc(kf_ctx* ctx, const KfLayout& L, double* acc, double* G, double* C, cudaStream_t st);
void kf_oz_destroy(kf_ctx* ctx);

// batch.cu
int kf_fit_batch_small(kf_ctx* ctx, int nprob, const kf_basis* const* bases, const kf_problem* probs, const kf_solve* solves, kf_result* outs,
                       const int* which, int nwhich, int* accepted);

// preprocess.cu: get_scale / get_zeta / get_snapshotPairs on the device
int kf_pp_scale(kf_ctx* ctx, const double* d_y, const double* d_u, long long T, int n, int m, double* d_scale, cudaStream_t st);
int kf_pp_count_pairs(kf_ctx* ctx, const double* d_t, long long T, int nd, long long* M_out, cudaStream_t st);
int kf_pp_pairs(kf_ctx* ctx, const double* d_t, const double* d_y, const double* d_u, const double* d_scale, long long T, int n, int m,
                int nd, long long M, double* d_alpha, double* d_beta, double* d_uo, cudaStream_t st);
int kf_fit_device_pairs(kf_ctx* ctx, const kf_problem* prob, const kf_solve* solve, kf_result* out, double t0, bool accumulated = false,
                        long long host_row0 = 0, long long host_ld = 0);
// api.cu: the fit of rows [lo, hi) of the host snapshot pairs on this context's device (kf_fit, kf_fit_multi)
int kf_fit_host_shard(kf_ctx* ctx, const kf_basis* basis, const kf_problem* prob, long long lo, long long hi, const kf_solve* solve,
                      kf_result* out);
// comm.cu: collectives on the context's NCCL communicator (no-ops without one), asynchronous on `st`
int kf_comm_allreduce(kf_ctx* ctx, double* buf, size_t count, int op_max, cudaStream_t st);
int kf_comm_allgather(kf_ctx* ctx, const double* send, double* recv, size_t count_per_rank, cudaStream_t st);
int kf_comm_allreduce_u64(kf_ctx* ctx, unsigned long long* buf, size_t count, cudaStream_t st);
int kf_comm_gather_blocks(kf_ctx* ctx, double* buf, const size_t* offs, const size_t* counts, cudaStream_t st);

// rollout.cu
int kf_rollout_impl(kf_ctx* ctx, int nmodels, const kf_model* mdls, int ntrials, const int* T, const double* const* zeta0,
                    const double* const* u, int nout, double* const* ysim);

// qp.cu
struct KfQpResult { double objective; double l1; int iters; double lam; int capped; double gap; double grad_inner; double grad_max; };
int kf_solve_l1ball_multi(kf_ctx* ctx, int P, int Pp, const double* G, const double* C, int nb, const double* t, int fix_c0,
                          int fix_c1, int max_iter, double tol, double* K_all, KfQpResult* res, cudaStream_t st);
// qp_as.cu: exact primal-dual active-set solver (per-column batched Cholesky); same contract as kf_solve_l1ball_multi
int kf_solve_l1ball_as(kf_ctx* ctx, int P, int Pp, const double* G, const double* C, int nb, const double* t, int fix_c0, int fix_c1,
                       int max_iter, double* K_all, KfQpResult* res, cudaStream_t st);
int kf_qp_scale_free(kf_ctx* ctx, double* K, int P, int Pp, int fix_c0, int fix_c1, double s, cudaStream_t st);
int kf_qp_deal_cols(kf_ctx* ctx, const double* src, double* dst, int P, int Pp, int R, int gather, cudaStream_t st);
int kf_add_diag(kf_ctx* ctx, double* G, int Pp, int P, double shift, cudaStream_t st);
int kf_qp_evaluate(kf_ctx* ctx, int P, int Pp, const double* G, const double* C, const double* K, int fix_c0, int fix_c1, double t_free,
                   KfQpResult* res, cudaStream_t st, int own_lo = 0, int own_hi = 0);
