/* koopfit.h — C ABI of libkoopfit.so: the B200-native Ksysid EDMD fit.
 *
 * The reference (roahmlab/koopman-realizations) is pure MATLAB and has NO FFI
 * for this path; the seam is created at Ksysid.get_Koopman.  Each entry point
 * below names the reference code it replaces (file:line into the reference
 * tree).  The MATLAB-side binding (a thin mexFunction) is shown in
 * INTEGRATION.md; in this repository the same ABI is driven through ctypes by
 * koopman-realizations_b200/ksysid.py.
 *
 * Conventions
 *   - all matrices are double, COLUMN-MAJOR (MATLAB layout); snapshot matrices
 *     alpha/beta (M x nzeta) and u (M x m) have leading dimension M, i.e. each
 *     variable is one contiguous column (coalesced on the device as-is).
 *   - caller owns every host buffer; the library owns device memory and streams
 *     inside kf_ctx.  No library-allocated pointer crosses the ABI except the
 *     device accumulator exposed by kf_accum_buffer (for the NCCL all-reduce).
 *   - every call returns 0 on success or a KF_E* code; kf_last_error gives text.
 *   - there is no CPU fallback: without a CUDA device kf_create fails.
 */
#ifndef KOOPFIT_H
#define KOOPFIT_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KF_VERSION 200

typedef struct kf_ctx kf_ctx;

enum { KF_OK = 0, KF_EINVAL = 1, KF_ECUDA = 2, KF_ENOMEM = 3, KF_ENUMERIC = 4, KF_EUNSUPPORTED = 5,
       KF_EAGAIN = 6,   /* staged API only: kf_solve_dev needs ANOTHER data pass over the shard (ill-conditioned regressor, see
                           kf_solve_dev); not an error */
       KF_ECOMM = 7 };  /* NCCL failure */

/* obj.model_type (Ksysid.m:96-104) */
enum { KF_LINEAR = 0, KF_BILINEAR = 1, KF_NONLINEAR = 2 };

/* obj.obs_type entries (Ksysid.m:486-501) */
enum { KF_POLY = 0, KF_FOURIER = 1, KF_FOURIER_SPARSER = 2, KF_GAUSSIAN = 3, KF_HERMITE = 4 };

/* one (obs_type{i}, obs_degree(i)) pair */
typedef struct {
    int type;               /* KF_POLY ... */
    int degree;             /* obs_degree(i) */
    const double* centres;  /* KF_GAUSSIAN: zeta0 (nv x degree, column-major; Ksysid.m:803); else NULL */
} kf_block;

/* the dictionary psi = [v; block_1; ...; 1] over v (Ksysid.m:484-505) */
typedef struct {
    int nv;                 /* length of v: nzeta (linear, bilinear) or nzeta+m (nonlinear; Ksysid.m:475-477) */
    int nblocks;
    const kf_block* blocks;
    const double* pcs;      /* dim_red: basis.pcs (N_full x n_pcs column-major, Ksysid.m:1507-1510) or NULL */
    int n_pcs;
} kf_basis;

/* snapshotPairs (Ksysid.m:1005) */
typedef struct {
    long long M;            /* number of snapshot pairs (rows) */
    int nzeta;              /* columns of alpha/beta */
    int m;                  /* columns of u */
    int model;              /* KF_LINEAR | KF_BILINEAR | KF_NONLINEAR */
    const double* alpha;    /* M x nzeta */
    const double* beta;     /* M x nzeta */
    const double* u;        /* M x m */
    int pc_cols;            /* opt-in fast mode: compute only the first pc_cols columns of K (and C); 0 = all P.
                               Downstream only K(:,1:N) (linear, bilinear: A, B — Ksysid.m:1199-1200, 1258-1259) or
                               K(:,1:nzeta) (nonlinear: F — 1329) is consumed; least-squares branch only. */
    int nw;                 /* `loaded` model (Ksysid.m:88-93, 539-626): columns of w; 0 = unloaded.  The lifted state becomes
                               [1; w] (x) psi  (psi, w_1 psi, ..., w_nw psi: Ksysid.m:594-599), so P = N (nw+1) [+ m | x (m+1)] */
    const double* w;        /* M x nw load of every snapshot pair (snapshotPairs.w, Ksysid.m:953-957), or NULL */
} kf_problem;

/* how to solve (Ksysid.m:1068-1080) */
enum { KF_LS_AUTO = 0, KF_LS_GRAM = 1, KF_LS_QR = 2 };
enum { KF_PSD_AS_REFERENCE = 0, KF_PSD_NEVER = 1, KF_PSD_ALWAYS = 2 };
typedef struct {
    int least_squares;      /* 1: K = Px \ Py (obj.lasso >= 1e6 branch, Ksysid.m:1068-1069) */
    int ls_method;          /* KF_LS_AUTO | KF_LS_GRAM (pivoted Cholesky of G) | KF_LS_QR (Householder QRCP of Px) */
    double pivot_tol;       /* KF_LS_GRAM relative pivot tolerance on |R_jj|/|R_11|; <=0 -> default 1e-7 */
    int nt;                 /* QP branch: number of L1 budgets */
    const double* t;        /* QP branch: t_i = lasso_i * N (Ksysid.m:996,1135) */
    int psd_shift;          /* KF_PSD_*: the `any(eig(G)<0)` +1e-6 I branch (Ksysid.m:1117-1120) */
    int delay_constraint;   /* 1: pin the delay columns (linear, nd>=1; Ksysid.m:1139-1164) */
    int n;                  /* params.n  (needed by the delay constraint) */
    int nd;                 /* params.nd */
    int qp_max_iter;        /* <=0 -> default.  Coordinate descent: sweeps per evaluation of the multiplier (100000 for P <= 256,
                               else 2000); active set: exact steps per attempt of a budget (200) */
    double qp_tol;          /* coordinate descent: relative iterate tolerance of a sweep; <=0 -> default 1e-13 (the active-set
                               solver is exact and ignores it) */
} kf_solve;

typedef struct {
    int rank;               /* numerical rank found by the LS solve */
    int ls_method_used;     /* KF_LS_GRAM or KF_LS_QR */
    int passes;             /* data passes (lift + contraction) */
    int psd_shift_applied;  /* QP: 1 if 1e-6 I was added */
    double min_pivot;       /* smallest accepted |R_jj| */
    double max_pivot;       /* |R_11| */
    double t_lift_gram_ms;  /* device time of lift + Gram */
    double t_solve_ms;      /* device time of the solve */
    double t_total_ms;      /* wall time of the call */
    int qp_capped;          /* QP: coordinate-descent evaluations that hit the sweep bound / active-set budgets that did not
                               settle (0 = all converged; kf_result.qp_gap tells how far off the returned K can be) */
    int reserved;
    double cond_est;        /* KF_LS_GRAM: max/min accepted |R_jj| of the first factorisation, a lower bound of cond(Px(:,basic));
                               above the "refine_kappa" option (1e3) the Gram route re-orthogonalises with extra data passes */
    int refine_passes;      /* KF_LS_GRAM: extra data passes of the multi-level pivoted Cholesky-QR refinement (0: one pass) */
    int refine_capped;      /* 1: the refinement hit its level limit before the basis was well conditioned (K is best effort) */
} kf_info;

/* caller-allocated outputs; NULL members are skipped */
typedef struct {
    double* K;              /* P x Pc x max(1,nt), Pc = pc_cols or P: koopData.K (Ksysid.m:1084) */
    double* G;              /* P x P : Px'Px (Ksysid.m:1114) */
    double* C;              /* P x P : Px'Py (Ksysid.m:1125) */
    double* Px;             /* M x P : regressor (Ksysid.m:1019-1065); koopData.Px = Px(:,1:N) */
    double* Py;             /* M x P */
    int* perm;              /* P: pivot order (0-based), first `rank` entries = basic set */
    double* objective;      /* nt: 0.5 tr(K'GK) - tr(C'K) per budget */
    double* l1norm;         /* nt: ||vec K||_1 */
    int* qp_iters;          /* nt: multiplier evaluations (coordinate descent) or exact active-set steps spent on the budget */
    double* qp_gap;         /* nt: certified optimality gap of the returned K: f(K) - min f <= qp_gap (Frank-Wolfe gap
                               <GK - C, K> + t ||GK - C||_inf over the free columns); compare with 1e-8 |objective| */
    kf_info info;
} kf_result;

/* ---- context ---------------------------------------------------------- */
int  kf_create(kf_ctx** ctx, int device);            /* one context per GPU / per process rank */
void kf_destroy(kf_ctx* ctx);
const char* kf_last_error(const kf_ctx* ctx);        /* ctx may be NULL for create-time errors */
int  kf_version(void);

/* ---- dictionary -------------------------------------------------------
 * Replaces def_observables / def_*Lift + matlabFunction (Ksysid.m:455-536, 629-863):
 * dimensions of the dictionary, N = params.N (Ksysid.m:534 / 1511-1517) and the
 * regressor width P (Ksysid.m:1019-1028). */
int kf_basis_dims(const kf_basis* basis, int model, int m, int* n_full, int* N, int* P);

/* exponent / multiplier table of one block in the reference's `partitions` order
 * (partitions.m:206-219; Ksysid.m:645-648, 747-751, 848-851).  rows*cols ints, row-major.
 * Call with table=NULL to query rows/cols. */
int kf_block_table(int type, int degree, int nv, int* rows, int* cols, int* table);

/* lift.econ_full evaluated on `rows` points (Ksysid.m:1443-1491, 1614-1618):
 * V (rows x nv) -> Psi (rows x N), both column-major HOST buffers; runs on the GPU. */
int kf_lift(kf_ctx* ctx, const kf_basis* basis, long long rows, const double* V, double* Psi);

/* Principal components of the lifted points: replaces `[coeffs, ~, ~, ~, explained] = pca(Psi)` in the dim_red branch of
 * get_econ_observables (Ksysid.m:1495-1517; lift_snapshots 1394-1432).  V (rows x nv, column-major HOST buffer) is lifted
 * through the FULL dictionary of `basis` (no pcs); mu (n_full) = feature means, latent (n_full) = eigenvalues of the sample
 * covariance in decreasing order, coeff (n_full x n_full, column-major) = principal directions with MATLAB's sign convention
 * (largest component of every column positive).  Everything runs on the GPU: two lift passes (means, then the Gram of the
 * CENTRED features on the DMMA GEMM) and a one-sided Jacobi eigensolver.  Any of mu / latent / coeff may be NULL. */
int kf_pca(kf_ctx* ctx, const kf_basis* basis, long long rows, const double* V, double* mu, double* latent, double* coeff);

/* ---- the fit ----------------------------------------------------------
 * Replaces Ksysid.get_Koopman (Ksysid.m:987-1092) + solve_KoopmanQP (1095-1176) and the
 * lasso loop of train_models (1370-1387): lifts once, accumulates G, C once, solves
 * LS or all nt budgets.  HOST pointers in `prob`; copies are inside the call. */
int kf_fit(kf_ctx* ctx, const kf_basis* basis, const kf_problem* prob, const kf_solve* solve, kf_result* out);

/* kf_fit with DEVICE pointers in `prob` (column-major, leading dimension M; this rank's shard on a multi-GPU context):
 * lift + Gram, the all-reduce over the context's communicator if it has one, the solve including any refinement passes.
 * Outputs go to HOST pointers as in kf_fit. */
int kf_fit_dev(kf_ctx* ctx, const kf_basis* basis, const kf_problem* prob, const kf_solve* solve, kf_result* out);

/* The same fit from the RAW merged training series: get_scale (Ksysid.m:180-229), get_zeta (868-907) and
 * get_snapshotPairs (910-984, snapshots = Inf) run on the device, so the series crosses the boundary once
 * (T (1 + n + m) doubles instead of M (2 nzeta + m)) and the pairs are written straight into the buffers the lift reads. */
typedef struct {
    long long T;            /* rows of the merged training trials (merge_trials, Ksysid.m:380-401) */
    int n, m, nd;           /* params.n, params.m, params.nd (delays) */
    int model;              /* KF_LINEAR | KF_BILINEAR | KF_NONLINEAR */
    const double* t;        /* T: time stamps; a pair is dropped where t does not increase (trial boundary, 948) */
    const double* y;        /* T x n column-major, unscaled */
    const double* u;        /* T x m */
    int prescaled;          /* 1: y, u are already scaled into [-1, 1] (offset 0, factor 1 are returned) */
    int pc_cols;            /* as kf_problem.pc_cols */
} kf_series;
typedef struct {
    double* y_offset;       /* n: params.scale.y_offset (Ksysid.m:226); NULL members are skipped */
    double* y_factor;       /* n */
    double* u_offset;       /* m */
    double* u_factor;       /* m */
    long long M;            /* out: snapshot pairs used = (kept pairs) - 1 (Ksysid.m:960) */
} kf_scale;
/* rows M of kf_result.Px / Py for a series, so the caller can allocate them (host scan of t) */
long long kf_series_pairs(long long T, int nd, const double* t);
int kf_fit_series(kf_ctx* ctx, const kf_basis* basis, const kf_series* series, const kf_solve* solve, kf_scale* scale, kf_result* out);

/* Many independent fits in one call (evaluate_rand_models.m:45-144 fits 23 small models per random system).
 * Problems with P <= 32 and no dim_red run CONCURRENTLY, one CTA per problem, entirely on chip (tile-wise Householder
 * TSQR of [Px | Py] + pivoted QR of the small triangular factor: mldivide semantics): all least-squares fits, and the QP
 * fits (no pinned columns, psd_shift != ALWAYS) whose LS solution has full rank and lies inside every budget — it is then
 * the minimiser of solve_KoopmanQP too (the script's nonlinear models use lasso = 4, usually inactive); objective, l1norm
 * and qp_gap are filled.  Every other problem is solved by kf_fit one after the other.  bases[i], probs[i], solves[i],
 * outs[i] describe problem i (HOST pointers, as in kf_fit); problems that share the same alpha pointer share one upload. */
int kf_fit_batch(kf_ctx* ctx, int nprob, const kf_basis* const* bases, const kf_problem* probs, const kf_solve* solves, kf_result* outs);

/* Generic MATLAB `A \ B` for a tall A (M x P) and B (M x Pc), HOST column-major buffers:
 * Householder QR with column pivoting on the GPU, rank by max(size(A))*eps(|R11|), basic
 * solution X (P x Pc).  Replaces `Mtranspose = L \ R` in get_model (Ksysid.m:1216).
 * perm (P ints) and rank may be NULL. */
int kf_mldivide(kf_ctx* ctx, long long M, int P, int Pc, const double* A, const double* B, double* X, int* perm, int* rank);

/* ---- validation rollouts -------------------------------
 * The fitted model as get_model / get_BLmodel / get_NLmodel return it (Ksysid.m:1179-1341), discrete time. */
typedef struct {
    int model;              /* KF_LINEAR | KF_BILINEAR | KF_NONLINEAR */
    int n, m, nzeta, N;     /* params.n, params.m, params.nzeta, params.N */
    const double* A;        /* N x N (linear, bilinear), column-major */
    const double* B;        /* N x m (linear) | N x (N m) = [B_1 ... B_m] (bilinear) */
    const double* F;        /* nonlinear: nzeta x N, F = K(:,1:nzeta)' (Ksysid.m:1329) */
} kf_model;

/* val_model / val_BLmodel / val_NLmodel (Ksysid.m:1623-1879; discrete, unloaded): open-loop simulation of `nmodels`
 * candidate models (e.g. the candidates of a lasso vector, train_models 1370-1387; all of one model type and size) on
 * `ntrials` validation trials at once, one CTA per (trial, candidate).  Trial k: T[k] samples, initial delay-embedded
 * state zeta0[k] (nzeta), inputs u[k] (T[k] x m column-major).  Output ysim[c * ntrials + k] (T[k] x nout column-major,
 * caller-allocated): the first nout state rows per step, nout = n gives y = C z with C = [I_n 0] (Ksysid.m:1203, 1262)
 * or zeta(1:n) (1864); nout = nzeta gives results.sim.zeta, nout = N results.sim.z (1700-1703); nout <= 0 means n.
 *   z_1 = lift.econ_full(zeta0); z+ = A z + B u (1685) | A z + Beta(z) u (1783) | zeta+ = F psi([zeta;u]) (1860).
 * Error metrics (get_error, 1882-1898) stay on the host. */
int kf_rollout(kf_ctx* ctx, const kf_basis* basis, int nmodels, const kf_model* models, int ntrials, const int* T,
               const double* const* zeta0, const double* const* u, int nout, double* const* ysim);

/* ---- consumer side: bilinear MPC cost assembly ------------------
 * Kmpc.get_costB_bilinear (Kmpc.m:569-596): the state-dependent condensed input matrix of the bilinear MPC problem, rebuilt at
 * every control step from the fitted model (A: N x N, B = [B_1 ... B_m]: N x N m as get_BLmodel returns them, Ksysid.m:1258-1259).
 *   block row i (i = 1..h) of the first block column = A^(i-1) Beta(z_i),  Beta(z) = B kron(I_m, z)  (Ksysid.m:1288-1289),
 *   z_i = z(i,:) when z has `horizon` rows, z(1,:) when it has one;  block column c = the first one shifted down by c-1 blocks.
 * nbatch problems in one launch: z is nbatch consecutive (nz x N) column-major matrices, Bout nbatch consecutive
 * (N (h+1)) x (m h) column-major matrices (HOST buffers). */
int kf_mpc_costB_bilinear(kf_ctx* ctx, int N, int m, int horizon, int nbatch, const double* A, const double* B, int nz,
                          const double* z, double* Bout);

/* ---- lasso sweep split across ranks (one process per GPU) -------------------
 * After the all-reduce of the partial Grams every rank holds the same G, C.  With a column partition the exact
 * active-set solver of rank r factors and updates only the columns [col_lo, col_hi) of K for ALL budgets; the columns
 * couple through the multiplier and the step control alone, i.e. a few doubles per step, which the library hands to
 * `allreduce` (op 0: sum, op 1: max over the ranks, in place; return 0 on success).  Every rank then takes the same
 * decisions; K comes back with only this rank's columns filled (the caller gathers the column blocks), objective, l1norm
 * and qp_gap are the global values.  col_hi <= col_lo restores the unpartitioned solve.  Not combinable with the pinned
 * delay columns (linear model with delays). */
typedef int (*kf_allreduce_fn)(void* user, double* vals, int n, int op);
int kf_set_qp_partition(kf_ctx* ctx, int col_lo, int col_hi, kf_allreduce_fn allreduce, void* user);

/* ---- staged, device-resident API (one rank of a snapshot-sharded fit) ---
 * kf_accumulate_dev: lift + Gram of this rank's shard; DEVICE pointers in `prob`;
 *   reset!=0 zeroes the accumulator first.  Asynchronous on the context stream.
 * kf_accum_buffer: the packed partial-Gram accumulator [G-tiles | C-tiles] to be
 *   summed across ranks (ncclAllReduce(sum, double) by the caller, e.g. torch.distributed).
 * kf_solve_dev: solve from the (reduced) accumulator; outputs to HOST pointers in `out`.
 *   Least squares on an ill-conditioned regressor (kf_info.cond_est above the "refine_kappa" option): the Gram route squares
 *   the condition number, so it re-orthogonalises with extra data passes (multi-level pivoted Cholesky-QR: every pass lifts the
 *   shard again, applies the basis change found so far and accumulates the Gram of the NEW features).  kf_solve_dev then returns
 *   KF_EAGAIN and the caller repeats, with the SAME shard(s):
 *       kf_accumulate_dev(ctx, basis, prob, 0);  [all-reduce kf_accum_buffer across ranks];  kf_solve_dev(ctx, solve, out);
 *   until it returns KF_OK (usually after one extra pass).  Every rank takes the same decisions (they all factor the same
 *   reduced matrices).  kf_fit / kf_fit_series / kf_fit_multi run this loop internally. */
int kf_accumulate_dev(kf_ctx* ctx, const kf_basis* basis, const kf_problem* prob, int reset);
/* Lift-only mode on DEVICE buffers (asynchronous on the context stream): the materialised regressors [Px | Py]
 * (Ksysid.m:1019-1065; M x 2P column-major, leading dimension ld >= M) of the device snapshot pairs in `prob`, and
 * lift.econ_full of `rows` device points (V rows x nv -> Psi rows x N).  HBM-bound: 8 (2 nzeta + m) B read and
 * 16 P B written per pair. */
int kf_regressors_dev(kf_ctx* ctx, const kf_basis* basis, const kf_problem* prob, double* dev_PxPy, long long ld);
int kf_lift_dev(kf_ctx* ctx, const kf_basis* basis, long long rows, const double* dev_V, double* dev_Psi);
int kf_accum_buffer(kf_ctx* ctx, double** dev_ptr, size_t* count);
int kf_solve_dev(kf_ctx* ctx, const kf_solve* solve, kf_result* out);
int kf_sync(kf_ctx* ctx);                            /* cudaStreamSynchronize(context stream) */
void* kf_stream(kf_ctx* ctx);                        /* the context's cudaStream_t */

/* ---- multi-GPU inside the library ---------------------------------------------------------
 * The reference's call site is ONE process (Ksysid.train_models -> get_Koopman, Ksysid.m:1357, 1373), so the library
 * owns the NCCL communicators (NCCL is resolved with dlopen at run time) and the MEX shim reaches every visible GPU
 * with one call:
 *
 *   kf_create_multi / kf_fit_multi   one process, ndev devices (device_ids NULL or ndev <= 0: all visible devices): a
 *       context and a host thread per device, ncclCommInitAll.  kf_fit_multi has kf_fit's contract on the FULL host
 *       snapshot pairs: rows are split into contiguous shards, every device copies its shard in blocks and lifts /
 *       contracts each block as it arrives, ONE ncclAllReduce sums the packed partial Grams (and the snapshot count) over
 *       NVLink, the solve is replicated (deterministic: every device takes the same rank / refinement decisions),
 *       device 0 writes K, G, C, perm and the per-budget outputs, every device its own rows of Px / Py.  A lasso vector
 *       solved by the exact active-set solver is split by COLUMNS of K over the devices (Ksysid.m:1370-1387: the
 *       reference re-fits per lasso value); the step scalars are reduced on the device in stream order.
 *   kf_comm_unique_id / kf_comm_init_rank   one process per GPU (torchrun / MPI / MATLAB parpool): rank 0 draws the
 *       128-byte id, the caller broadcasts it, every rank's context joins; kf_fit on such a context takes THIS RANK's
 *       shard of the snapshot pairs and behaves as above (every rank receives the full result).
 * KF_LS_QR is refused on a multi-GPU fit (it needs all regressors on one device); KF_LS_AUTO takes the Gram route, which
 * re-orthogonalises by itself on ill-conditioned data (kf_info.refine_passes). */
typedef struct kf_multi kf_multi;
int  kf_create_multi(kf_multi** mc, const int* device_ids, int ndev);
void kf_destroy_multi(kf_multi* mc);
int  kf_multi_size(const kf_multi* mc);
kf_ctx* kf_multi_ctx(kf_multi* mc, int i);                 /* device i's context (kf_counters, kf_last_times, ...) */
const char* kf_multi_last_error(const kf_multi* mc);
int  kf_multi_set_option(kf_multi* mc, const char* name, double value);   /* kf_set_option on every device */
int  kf_fit_multi(kf_multi* mc, const kf_basis* basis, const kf_problem* prob, const kf_solve* solve, kf_result* out);
#define KF_COMM_ID_BYTES 128
int  kf_comm_unique_id(void* id, size_t bytes);            /* ncclGetUniqueId (rank 0) */
int  kf_comm_init_rank(kf_ctx* ctx, int nranks, int rank, const void* id, size_t bytes);   /* ncclCommInitRank on ctx's device */
int  kf_comm_destroy(kf_ctx* ctx);                         /* also done by kf_destroy */
int  kf_comm_info(const kf_ctx* ctx, int* nranks, int* rank, int* nccl_version);

/* ---- instrumentation for bench.py / tests ----------------------------- */
/* flops issued to the FP64 tensor pipe and kernel launches since the last reset */
int kf_counters(kf_ctx* ctx, double* dmma_flops, long long* launches, int reset);
/* which Gram engine the last accumulation used (1: FP64 DMMA, 2: INT8 tensor cores / Ozaki scheme II — FP64-exact), the INT8
 * tensor operations issued since the last reset and per contraction launch, the contraction launches of the last accumulation
 * and how many of them were timed in isolation ("profile" option) */
int kf_engine_info(kf_ctx* ctx, int* engine, double* i8_ops, double* i8_ops_per_launch, long long* gram_launches, int* gram_sampled, int reset);
/* device time (ms, CUDA events on the context stream) of the last lift+Gram phase and solve phase; gram_kernel_ms is an ESTIMATE:
 * mean duration of the isolated, sampled contraction launches x the number of launches */
int kf_last_times(kf_ctx* ctx, double* lift_gram_ms, double* gram_kernel_ms, double* solve_ms);
/* tuning knobs; returns KF_EINVAL if unknown:
 *   "chunk" (snapshots per L2-resident panel), "panel_mb", "splitk", "overlap" (two chunk pipelines), "tma" (Gram operands by
 *   tensor-map TMA = 1 / cp.async = 0), "profile" (sample Gram-kernel durations), "qr_max_gb" (KF_LS_AUTO takes the QRCP route
 *   up to this size of [Px | Py]), "qp_method" (0 auto = 2: exact active set; 1: coordinate descent in lockstep; 2),
 *   "as_frac" (active set: bound on the pattern change per step, fraction of the support, default 0.05, self-tuning downwards),
 *   "as_ws_gb" (active set: bound on the factor workspace), "lift_tile" (materialising lift: shared-memory tile kernel = 1),
 *   "refine" (Gram-route refinement: 0 off, 1 adaptive = default, 2 always at least one extra pass), "refine_kappa" (pivot-ratio
 *   threshold, default 1e3), "refine_level_tol" (dynamic range one level resolves, default 1e-5), "refine_max" (level limit, 4),
 *   "qp_split" (multi-GPU context: split the active-set lasso sweep by columns over the ranks, default 1),
 *   "gram_engine" (0 auto: INT8 tensor cores for P >= 1024 and >= 4 panels of snapshots, FP64 DMMA otherwise; 1 DMMA; 2 INT8),
 *   "oz_sym" (INT8 engine: exploit the Kronecker block symmetry of a bilinear regressor, default 1),
 *   "lift_wide" (materialising lift / panel lift: streaming kernel with 32- or 64-snapshot tiles = 1, default), "lift_ls" (force the
 *   tile width: 32 | 64; 8 | 16 for the narrow fallback kernel), "lift_smem_kb" (shared memory per CTA of the streaming kernel,
 *   default 110), "lift_minb" (2 | 3 resident CTAs per SM it is compiled for), "lift_panel_fit" (fit path: panel lift through the
 *   streaming evaluator, default 1),
 *   "qr_blocked" (QRCP route: 1 = blocked Householder QR of the tall matrix, then pivoted QR of R1 = default; 0 = column by column;
 *   2 = diagnostic: two-stage without compact WY), "qr_nb" (panel width 32 | 64 | 128, default 64), "qr_ksplit" (chunks of the
 *   two-level summation over the rows, default 64),
 *   "as_skip" (active set: columns whose support did not change are not factored again, default 1), "as_level" (level-synchronous
 *   factorisation: -1 off = default, 0 auto, 1 always), "as_diag" (print per-step diagnostics to stderr) */
int kf_set_option(kf_ctx* ctx, const char* name, double value);

#ifdef __cplusplus
}
#endif
#endif /* KOOPFIT_H */
