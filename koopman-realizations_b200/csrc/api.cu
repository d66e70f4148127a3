// C-ABI entry points of libkoopfit.so (see include/koopfit.h for the contract and the
// reference lines each call replaces).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "kf_internal.h"
#include "tma_host.h"

namespace {

thread_local std::string g_create_err;

}  // namespace
std::mutex& kf_smem_table_mutex() { static std::mutex m; return m; }
std::map<std::pair<int, const void*>, size_t>& kf_smem_table() { static std::map<std::pair<int, const void*>, size_t> t; return t; }
namespace {
double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int pair_index(int a, int b, int m) { return a * (m + 1) - a * (a - 1) / 2 + (b - a); }

// compile + upload the dictionary
int prepare_program(kf_ctx* ctx, const kf_basis* basis) {
    std::string err;
    int rc = kf_build_program(basis, ctx->prog, err);
    if (rc) {
        ctx->err = err;
        return rc;
    }
    const KfProgram& p = ctx->prog;
    {   // the lift-group tables on the device are keyed by the dictionary: the same dictionary again (a loop of fits) keeps them
        unsigned long long h = 1469598103934665603ull;
        auto mix = [&](const void* data, size_t nbytes) {
            const unsigned char* b = static_cast<const unsigned char*>(data);
            for (size_t i = 0; i < nbytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
        };
        const int dims[3] = {p.nv, p.n_pcs, p.ngauss};
        mix(dims, sizeof(dims));
        mix(p.ops.data(), p.ops.size() * sizeof(KfOp));
        mix(p.centres.data(), p.centres.size() * sizeof(double));
        mix(p.pcs.data(), p.pcs.size() * sizeof(double));
        if (h != ctx->prog_hash || ctx->prog_gen == 0) { ctx->prog_hash = h; ctx->prog_gen += 1; }
    }
    KF_CUDA(ctx, ctx->d_ops.ensure(sizeof(KfOp) * p.ops.size()));
    KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_ops.p, p.ops.data(), sizeof(KfOp) * p.ops.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!p.centres.empty()) {
        KF_CUDA(ctx, ctx->d_centres.ensure(sizeof(double) * p.centres.size()));
        KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_centres.p, p.centres.data(), sizeof(double) * p.centres.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (!p.pcs.empty()) {
        KF_CUDA(ctx, ctx->d_pcs.ensure(sizeof(double) * p.pcs.size()));
        KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_pcs.p, p.pcs.data(), sizeof(double) * p.pcs.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    std::vector<int> order;
    kf_program_levels(p, order, ctx->level_start);
    KF_CUDA(ctx, ctx->d_order.ensure(sizeof(int) * order.size()));
    KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_order.p, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice, ctx->stream));
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // host vectors may be rebuilt by the next call
    return KF_OK;
}

int check_problem(kf_ctx* ctx, const kf_problem* pr) {
    if (!pr || pr->M <= 0 || pr->nzeta <= 0 || pr->m < 0 || !pr->alpha || !pr->beta || (pr->m > 0 && !pr->u)) {
        ctx->err = "kf_problem: M, nzeta must be > 0 and alpha/beta/u non-NULL";
        return KF_EINVAL;
    }
    if (pr->model != KF_LINEAR && pr->model != KF_BILINEAR && pr->model != KF_NONLINEAR) {
        ctx->err = "Invalid model_type chosen. Must be linear, bilinear, or nonlinear.";   // Ksysid.m:103
        return KF_EINVAL;
    }
    if (pr->nw < 0 || (pr->nw > 0 && !pr->w)) {
        ctx->err = "kf_problem: a loaded model (nw > 0) needs the loads w (M x nw)";
        return KF_EINVAL;
    }
    if (pr->nw > 0 && pr->pc_cols > 0) {
        ctx->err = "kf_problem: pc_cols is not supported for loaded models";
        return KF_EINVAL;
    }
    const int nv = pr->nzeta + (pr->model == KF_NONLINEAR ? pr->m : 0);
    if (ctx->prog.nv != nv) {
        ctx->err = "kf_basis.nv must equal nzeta (linear, bilinear) or nzeta+m (nonlinear)";
        return KF_EINVAL;
    }
    return KF_OK;
}

// layout of the panel, the tile list and the accumulator for (program, model, m, M)
int make_layout(kf_ctx* ctx, const kf_problem* pr) {
    KfLayout L;
    const KfProgram& p = ctx->prog;
    L.model = pr->model;
    L.m = pr->m;
    L.nzeta = pr->nzeta;
    L.nv = p.nv;
    L.n_full = p.n_full();
    L.N = p.N();
    L.nw = pr->nw;
    L.dense = pr->nw > 0;
    L.P = kf_regressor_width(L.model, L.N, L.m, L.nw);
    L.Rx = L.N + (L.model == KF_LINEAR ? L.m : 0);
    L.Rxp = (int)kf_roundup(L.Rx, KF_BM);
    L.Ny = L.N;
    L.Nyp = (int)kf_roundup(L.N, KF_BN);
    L.nW = (L.model == KF_BILINEAR) ? (L.m + 1) * (L.m + 2) / 2 : 0;
    L.x_off = 0;
    L.y_off = L.Rxp;
    L.w_off = L.Rxp + L.Nyp;
    L.rows = L.w_off + (int)kf_roundup(L.nW, 8);
    L.Pp = (int)kf_roundup(L.P, KF_BM);
    L.Pc = (pr->pc_cols > 0 && pr->pc_cols < L.P) ? pr->pc_cols : L.P;
    if (L.dense) {
        // `loaded` model: [1; u] (x) [1; w] (x) psi.  The weighted-block accumulator is specialised to [1; u] (x) psi, so G and C are
        // accumulated densely from the materialised regressors of each chunk (dense_pass) — same collective, same solvers.
        L.slab = 2LL * L.Pp * L.Pp + KF_ACC_TRAILER;
        long long Mc = ctx->opt_chunk > 0 ? ctx->opt_chunk : (long long)(ctx->opt_panel_mb * 1048576.0 / (8.0 * (2.0 * L.P + KF_BM)));
        Mc = std::max<long long>(256, std::min<long long>(Mc / 256 * 256, 8192));
        L.Mc = (int)Mc;
        L.nsplit = 1;
        L.valid = true;
        KF_CUDA(ctx, ctx->rf.d_dense.ensure((size_t)L.slab * sizeof(double)));
        ctx->lay = L;
        return KF_OK;
    }

    // chunk: one panel should stay L2-resident
    long long Mc = ctx->opt_chunk > 0 ? ctx->opt_chunk
                                      : (long long)(ctx->opt_panel_mb * 1048576.0 / (8.0 * L.rows));
    Mc = std::max<long long>(256, std::min<long long>(Mc, 32768));
    Mc = std::min<long long>(Mc, kf_roundup(pr->M, 256));
    Mc = kf_roundup(Mc, 256);
    L.Mc = (int)Mc;

    // Gram engine: the INT8 tensor cores (Ozaki scheme II, ozaki.cu) for the large contractions, the FP64 DMMA kernel otherwise
    if ((ctx->opt_gram_engine == 2 || (ctx->opt_gram_engine == 0 && L.P >= 1024 && pr->M >= 4LL * L.Mc)) && kf_oz_supported(L)) {
        L.oz = true;
        L.dense = true;
        if (ctx->opt_chunk <= 0) {
            // the CRT / accumulate pass costs the same per panel whatever its depth, so deeper panels amortise it: ~100 MB of lifted
            // panel (6144 snapshots at config 5) measured +2.8 % over 64 MB; 8192 no better (lower clocks at the power cap)
            long long McOz = (long long)(100.0 * 1048576.0 / (8.0 * L.rows)) / 256 * 256;
            McOz = std::max<long long>(1024, std::min<long long>(McOz, 8192));
            L.Mc = (int)std::max<long long>(L.Mc, std::min<long long>(McOz, kf_roundup(pr->M, 256)));
        }
        L.slab = 2LL * L.Pp * L.Pp + KF_ACC_TRAILER;
        L.nsplit = 1;
        // chunk pipelines in flight (option "oz_pipes")
        L.npipes = std::max(2, std::min(ctx->opt_oz_pipes, KF_MAX_PIPES));
        const size_t panel_bytes = (size_t)L.rows * L.Mc * sizeof(double);
        for (int b = 0; b < L.npipes; ++b) {
            KF_CUDA(ctx, ctx->d_panel[b].ensure(panel_bytes));
            KF_CUDA(ctx, cudaMemsetAsync(ctx->d_panel[b].p, 0, panel_bytes, ctx->stream));
        }
        if (p.n_pcs) KF_CUDA(ctx, ctx->d_full.ensure((size_t)L.npipes * 2 * L.n_full * L.Mc * sizeof(double)));   // dim_red scratch per pipeline
        KF_CUDA(ctx, ctx->rf.d_dense.ensure((size_t)L.slab * sizeof(double)));
        KF_CUDA(ctx, ctx->rf.d_dense2.ensure((size_t)L.slab * sizeof(double)));
        for (int b = 2; b < L.npipes; ++b) KF_CUDA(ctx, ctx->rf.d_dense_x[b - 2].ensure((size_t)L.slab * sizeof(double)));
        KF_TRY(kf_oz_prepare(ctx, L));
        L.valid = true;
        ctx->lay = L;
        return KF_OK;
    }

    const int tmx = L.Rxp / KF_BM, tny = L.Nyp / KF_BN;
    for (int a = 0; a <= (L.model == KF_BILINEAR ? L.m : 0); ++a)
        for (int b = a; b <= (L.model == KF_BILINEAR ? L.m : 0); ++b) {
            const int q = (L.model == KF_BILINEAR) ? pair_index(a, b, L.m) : 0;
            for (int tm = 0; tm < tmx; ++tm)
                for (int tn = 0; tn <= tm; ++tn) L.tiles.push_back(KfTile{0, q, a, b, tm, tn});
            // C tiles: only the column blocks that reach the first Pc columns of K (pc_cols fast mode).
            // bilinear: stored block (a,b), a <= b, serves C[(a,.),(b,.)] and C[(b,.),(a,.)] -> needed iff a < ceil(Pc/N);
            // otherwise: tile columns below min(Pc, N) (the u columns of a linear model come from G).
            int tn_end = tny;
            if (L.model == KF_BILINEAR) {
                if (a >= (L.Pc + L.N - 1) / L.N) tn_end = 0;
            } else {
                tn_end = std::min(tny, (std::min(L.Pc, L.N) + KF_BN - 1) / KF_BN);
            }
            for (int tm = 0; tm < tmx; ++tm)
                for (int tn = 0; tn < tn_end; ++tn) L.tiles.push_back(KfTile{1, q, a, b, tm, tn});
        }
    const int T = (int)L.tiles.size();
    L.slab = (long long)T * KF_TILE_ELEMS + KF_ACC_TRAILER;
    const int ksteps = L.Mc / KF_BK;
    // split-K so that small problems still fill the machine: aim at ~2 waves of 2 CTAs per SM
    const int ctas = T * (KF_BM / KF_CTA_M) * (KF_BN / KF_CTA_N);
    int nsplit = ctx->opt_splitk > 0 ? ctx->opt_splitk : (4 * ctx->sm_count + ctas - 1) / ctas;
    nsplit = std::max(1, std::min(nsplit, std::max(1, ksteps / 4)));
    L.nsplit = nsplit;
    L.valid = true;

    // buffers
    const size_t panel_bytes = (size_t)L.rows * L.Mc * sizeof(double);
    for (int b = 0; b < 2; ++b) {
        KF_CUDA(ctx, ctx->d_panel[b].ensure(panel_bytes));
        KF_CUDA(ctx, cudaMemsetAsync(ctx->d_panel[b].p, 0, panel_bytes, ctx->stream));   // pad rows must be finite
    }
    if (p.n_pcs) KF_CUDA(ctx, ctx->d_full.ensure((size_t)2 * 2 * L.n_full * L.Mc * sizeof(double)));   // dim_red scratch, one per chunk pipeline
    // two slab sets (one per chunk pipeline) x nsplit split-K slabs; folded into slab 0 by kf_reduce_slabs
    KF_CUDA(ctx, ctx->d_accum.ensure((size_t)2 * nsplit * L.slab * sizeof(double)));
    KF_CUDA(ctx, ctx->d_tilemeta.ensure(sizeof(KfTile) * T));
    KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_tilemeta.p, L.tiles.data(), sizeof(KfTile) * T, cudaMemcpyHostToDevice, ctx->stream));

    // task lists, one per panel buffer
    const int per = (ksteps + nsplit - 1) / nsplit;
    std::vector<KfGemmTask> tasks;
    std::vector<KfTmaTask> ttasks;
    for (int b = 0; b < 2; ++b) {
        tasks.clear();
        ttasks.clear();
        double* panel = ctx->d_panel[b].as<double>();
        for (int s = 0; s < nsplit; ++s) {
            const int k0 = std::min(s * per, ksteps) * KF_BK, k1 = std::min((s + 1) * per, ksteps) * KF_BK;
            if (k1 <= k0) continue;
            // every 128 x 128 accumulator tile is covered by (128/CTA_M) x (128/CTA_N) CTA tasks
            for (int t = 0; t < T; ++t)
                for (int sm = 0; sm < KF_BM / KF_CTA_M; ++sm)
                    for (int sn = 0; sn < KF_BN / KF_CTA_N; ++sn) {
                        const KfTile& tl = L.tiles[t];
                        KfGemmTask g{};
                        g.A = panel + (long long)(L.x_off + tl.tm * KF_BM + sm * KF_CTA_M) * L.Mc;
                        g.B = panel + (long long)((tl.kind == 0 ? L.x_off : L.y_off) + tl.tn * KF_BN + sn * KF_CTA_N) * L.Mc;
                        g.W = tl.q > 0 ? panel + (long long)(L.w_off + tl.q) * L.Mc : nullptr;
                        g.out = ctx->d_accum.as<double>() + (long long)(b * nsplit + s) * L.slab + (long long)t * KF_TILE_ELEMS +
                                (long long)sm * KF_CTA_M * KF_BN + sn * KF_CTA_N;
                        g.lda = g.ldb = L.Mc;
                        g.ldm = KF_BN;
                        g.ldn = 1;
                        g.k0 = k0;
                        g.k1 = k1;
                        g.a_rows = KF_CTA_M;
                        g.b_rows = KF_CTA_N;
                        g.alpha = 1.0;
                        g.accumulate = 1;
                        tasks.push_back(g);
                        KfTmaTask tt{};   // same CTA task for the tensor-map TMA kernel (128 x 64 sub-tile)
                        tt.a_row = L.x_off + tl.tm * KF_BM + sm * KF_CTA_M;
                        tt.b_row = (tl.kind == 0 ? L.x_off : L.y_off) + tl.tn * KF_BN + sn * KF_CTA_N;
                        tt.w_row = tl.q > 0 ? L.w_off + tl.q : -1;
                        tt.k0 = k0;
                        tt.k1 = k1;
                        tt.out = g.out;
                        tt.W = g.W;
                        ttasks.push_back(tt);
                    }
        }
        KF_CUDA(ctx, ctx->d_tasks[b].ensure(sizeof(KfGemmTask) * tasks.size()));
        KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_tasks[b].p, tasks.data(), sizeof(KfGemmTask) * tasks.size(), cudaMemcpyHostToDevice, ctx->stream));
        KF_CUDA(ctx, ctx->d_tma_tasks[b].ensure(sizeof(KfTmaTask) * ttasks.size()));
        KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_tma_tasks[b].p, ttasks.data(), sizeof(KfTmaTask) * ttasks.size(), cudaMemcpyHostToDevice, ctx->stream));
        if (ctx->opt_tma) {
            const int rc = kf_make_panel_tensor_map(&ctx->tmap[b], panel, (unsigned long long)L.Mc, (unsigned long long)L.rows);
            if (rc) {
                ctx->err = "cuTensorMapEncodeTiled failed for the lifted panel (rc=" + std::to_string(rc) + ")";
                return KF_ECUDA;
            }
        }
        KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->lay = L;
    return KF_OK;
}

int ntasks_of(const KfLayout& L) {
    const int ksteps = L.Mc / KF_BK;
    const int per = (ksteps + L.nsplit - 1) / L.nsplit;
    int used = 0;
    for (int s = 0; s < L.nsplit; ++s)
        if (std::min((s + 1) * per, ksteps) > std::min(s * per, ksteps)) ++used;
    return used * (int)L.tiles.size() * (KF_BM / KF_CTA_M) * (KF_BN / KF_CTA_N);
}

bool same_layout(const KfLayout& L, const KfProgram& p, const kf_problem* pr, int Mc_hint) {
    (void)Mc_hint;
    const int P = kf_regressor_width(pr->model, p.N(), pr->m, pr->nw);
    const int Pc = (pr->pc_cols > 0 && pr->pc_cols < P) ? pr->pc_cols : P;
    return L.valid && L.model == pr->model && L.m == pr->m && L.nzeta == pr->nzeta && L.n_full == p.n_full() &&
           L.N == p.N() && L.Pc == Pc && L.nw == pr->nw;
}

int dense_pass(kf_ctx* ctx, const kf_problem* pr, const double* St, double* GC, int Mc, long long c0, long long c1);
double* accum_base(kf_ctx* ctx) { return ctx->lay.dense ? ctx->rf.d_dense.as<double>() : ctx->d_accum.as<double>(); }

// lift + Gram of one shard whose snapshots are on the device; chunks [c0, c1) of the shard only (c1 < 0: to the end), so that
// a host shard can be accumulated block by block while the next block is still being copied in (fit_host_shard)
int accumulate_dev(kf_ctx* ctx, const kf_problem* pr, bool reset, long long c0 = 0, long long c1 = -1) {
    KF_TRY(check_problem(ctx, pr));
    if (reset || !same_layout(ctx->lay, ctx->prog, pr, 0)) {
        if (!reset && ctx->lay.valid) {
            ctx->err = "kf_accumulate_dev: layout changed without reset";
            return KF_EINVAL;
        }
        KF_TRY(make_layout(ctx, pr));
        const KfLayout& L = ctx->lay;
        if (L.oz) {
            KF_CUDA(ctx, cudaMemsetAsync(ctx->rf.d_dense2.p, 0, (size_t)L.slab * sizeof(double), ctx->stream));
            for (int b = 2; b < L.npipes; ++b) KF_CUDA(ctx, cudaMemsetAsync(ctx->rf.d_dense_x[b - 2].p, 0, (size_t)L.slab * sizeof(double), ctx->stream));
        }
        if (L.dense) KF_CUDA(ctx, cudaMemsetAsync(ctx->rf.d_dense.p, 0, (size_t)L.slab * sizeof(double), ctx->stream));
        else KF_CUDA(ctx, cudaMemsetAsync(ctx->d_accum.p, 0, (size_t)2 * L.nsplit * L.slab * sizeof(double), ctx->stream));
        ctx->accum_M = 0;
        ctx->rf.pending = false;
    }
    const KfLayout& L = ctx->lay;
    const KfProgram& p = ctx->prog;
    if (L.dense && !L.oz) {
        const long long nch = (pr->M + L.Mc - 1) / L.Mc;
        if (c1 < 0 || c1 > nch) c1 = nch;
        if (c0 == 0) KF_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
        KF_TRY(dense_pass(ctx, pr, nullptr, ctx->rf.d_dense.as<double>(), L.Mc, c0, c1));
        KF_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
        if (c1 == nch) ctx->accum_M += pr->M;
        ctx->last_gram_kernel_ms = 0.f;
        return KF_OK;
    }
    const int ntasks = L.oz ? 0 : ntasks_of(L);
    const bool weighted = (L.model == KF_BILINEAR);
    const long long nchunks = (pr->M + L.Mc - 1) / L.Mc;
    if (c1 < 0 || c1 > nchunks) c1 = nchunks;
    const bool first = (c0 == 0), last = (c1 == nchunks);
    // Two software pipelines: even chunks run lift -> Gram on stream A (panel 0, slab set 0), odd chunks on
    // stream B (panel 1, slab set 1).  The pipelines are independent (own panel, own accumulator slabs), so the
    // tail wave / epilogue of one Gram launch is filled by the next chunk's CTAs and the lifts hide under DMMAs.
    const bool overlap = ctx->opt_overlap && nchunks > 1;
    const int NP = overlap ? (L.oz ? L.npipes : 2) : 1;
    cudaStream_t S[KF_MAX_PIPES] = {ctx->stream, overlap ? ctx->stream2 : ctx->stream, ctx->stream_x[0], ctx->stream_x[1]};

    if (first) {
        KF_CUDA(ctx, cudaEventRecord(ctx->ev[0], S[0]));
        ctx->gram_ms_sampled = 0.f;
        ctx->gram_samples = 0;
    }
    if (overlap) {   // fork: pipeline B starts behind everything the context stream has been told to wait for
        KF_CUDA(ctx, cudaEventRecord(ctx->ev_fork, S[0]));
        for (int q = 1; q < NP; ++q) KF_CUDA(ctx, cudaStreamWaitEvent(S[q], ctx->ev_fork, 0));
    }

    for (long long c = c0; c < c1; ++c) {
        const int b = (int)(c % NP);
        cudaStream_t sl = S[b], sg = S[b];
        KfLiftArgs a{};
        a.ops = ctx->d_ops.as<KfOp>();
        a.centres = ctx->d_centres.as<double>();
        a.pcs = ctx->d_pcs.as<double>();
        a.order = ctx->d_order.as<int>(); a.nsides = 2; a.extras = 1;
        a.nv = p.nv; a.n_full = p.n_full(); a.n_pcs = p.n_pcs; a.N = L.N;
        a.nzeta = L.nzeta; a.m = L.m; a.model = L.model;
        a.alpha = pr->alpha; a.beta = pr->beta; a.u = pr->u;
        a.M = pr->M; a.start = c * L.Mc; a.Mc = L.Mc;
        a.panel = ctx->d_panel[b].as<double>(); a.ld = L.Mc;
        a.full = ctx->d_full.as<double>() + (p.n_pcs ? (size_t)b * 2 * L.n_full * L.Mc : 0);   // own scratch: the pipelines lift concurrently
        a.x_off = L.x_off; a.y_off = L.y_off; a.w_off = L.w_off; a.nW = L.nW;
        {
            int lrc = KF_OK;
            if (!(ctx->opt_lift_panel_fit && kf_launch_lift_panel_tile(ctx, a, sl, &lrc))) KF_TRY(kf_launch_lift(ctx, a, sl));
            else if (lrc) return lrc;
        }
        // CUDA-event timing of the Gram kernel on the stream it is launched on.  A sampled launch is isolated
        // from the other pipeline (which waits), so the duration is the kernel's own, not a time-shared one.
        const bool sample = ctx->opt_profile && (nchunks < 8 || (c % 61) == 3);
        if (sample) {
            if (overlap) {       // isolate the sampled launch: wait for every other pipeline
                for (int q = 0; q < NP; ++q)
                    if (q != b) {
                        KF_CUDA(ctx, cudaEventRecord(ctx->ev_join[q], S[q]));
                        KF_CUDA(ctx, cudaStreamWaitEvent(sg, ctx->ev_join[q], 0));
                    }
            }
            if (!L.oz) KF_CUDA(ctx, cudaEventRecord(ctx->ev[2], sg));
        }
        if (L.oz) {
            double* acc = (b == 0 ? ctx->rf.d_dense : b == 1 ? ctx->rf.d_dense2 : ctx->rf.d_dense_x[b - 2]).as<double>();
            KF_TRY(kf_oz_chunk(ctx, L, b, ctx->d_panel[b].as<double>(), acc, acc + (size_t)L.Pp * L.Pp, sg, sample ? ctx->ev[2] : nullptr,
                               sample ? ctx->ev[8] : nullptr));
        } else if (ctx->opt_tma)
            KF_TRY(kf_launch_gram_tma(ctx, ctx->d_tma_tasks[b].as<KfTmaTask>(), ntasks, weighted, ctx->tmap[b], sg));
        else
            KF_TRY(kf_launch_gemm_tasks(ctx, ctx->d_tasks[b].as<KfGemmTask>(), ntasks, weighted, sg));
        if (sample) {
            KF_CUDA(ctx, cudaEventRecord(ctx->ev[3], sg));
            if (overlap)
                for (int q = 0; q < NP; ++q)
                    if (q != b) KF_CUDA(ctx, cudaStreamWaitEvent(S[q], ctx->ev[3], 0));
            KF_CUDA(ctx, cudaEventSynchronize(ctx->ev[3]));
            float ms = 0.f;
            KF_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[2], L.oz ? ctx->ev[8] : ctx->ev[3]));
            ctx->gram_ms_sampled += ms;
            ++ctx->gram_samples;
        }
        if (!L.oz) ctx->dmma_flops += (double)L.tiles.size() * 2.0 * KF_TILE_ELEMS * (double)L.Mc;
    }
    if (overlap) {   // join the other pipelines into the context stream
        for (int q = 1; q < NP; ++q) {
            KF_CUDA(ctx, cudaEventRecord(ctx->ev_join[q], S[q]));
            KF_CUDA(ctx, cudaStreamWaitEvent(S[0], ctx->ev_join[q], 0));
        }
    }
    KF_CUDA(ctx, cudaEventRecord(ctx->ev[1], S[0]));
    if (last) {
        ctx->accum_M += pr->M;
        // an ESTIMATE: mean duration of the sampled (isolated) Gram launches x the number of launches
        ctx->last_gram_kernel_ms = ctx->gram_samples ? ctx->gram_ms_sampled / ctx->gram_samples * (float)nchunks : 0.f;
        ctx->last_gram_launches = nchunks;
        ctx->last_engine = L.oz ? 2 : 1;
        ctx->last_gram_sampled = ctx->gram_samples;
    }
    return KF_OK;
}

int finish_accum(kf_ctx* ctx) {
    const KfLayout& L = ctx->lay;
    if (L.oz) {      // fold the second pipeline's accumulators into the first (and clear them: idempotent)
        KF_TRY(kf_oz_finish(ctx, L, ctx->rf.d_dense.as<double>(), ctx->rf.d_dense2.as<double>(), ctx->stream));
        for (int b = 2; b < L.npipes; ++b)
            KF_TRY(kf_oz_finish(ctx, L, ctx->rf.d_dense.as<double>(), ctx->rf.d_dense_x[b - 2].as<double>(), ctx->stream));
        return KF_OK;
    }
    if (L.dense) return KF_OK;
    KF_TRY(kf_reduce_slabs(ctx, ctx->d_accum.as<double>(), L.slab, 2 * L.nsplit, ctx->stream));
    return KF_OK;
}

// the snapshot count of this rank into the trailer of slab 0: the all-reduce of the packed accumulator then also sums it, so
// every rank knows the total M (MATLAB's rank tolerance max(size(Px)) * eps(|R11|) needs it) without a second collective
int write_trailer(kf_ctx* ctx) {
    const KfLayout& L = ctx->lay;
    ctx->accum_M_d = (double)ctx->accum_M;
    KF_CUDA(ctx, cudaMemcpyAsync(accum_base(ctx) + (L.slab - KF_ACC_TRAILER), &ctx->accum_M_d, sizeof(double),
                                 cudaMemcpyHostToDevice, ctx->stream));
    return KF_OK;
}

int copy_out_matrix(kf_ctx* ctx, const double* d, int Pp, int P, double* h, int ncols = -1) {
    if (!h) return KF_OK;
    KF_CUDA(ctx, cudaMemcpy2DAsync(h, (size_t)P * sizeof(double), d, (size_t)Pp * sizeof(double), (size_t)P * sizeof(double),
                                   ncols < 0 ? P : ncols, cudaMemcpyDeviceToHost, ctx->stream));
    return KF_OK;
}

KfLiftArgs lift_args_of(kf_ctx* ctx, const kf_problem* pr) {
    KfLiftArgs a{};
    const KfProgram& p = ctx->prog;
    a.ops = ctx->d_ops.as<KfOp>(); a.centres = ctx->d_centres.as<double>(); a.pcs = ctx->d_pcs.as<double>();
    a.order = ctx->d_order.as<int>();
    a.nv = p.nv; a.n_full = p.n_full(); a.n_pcs = p.n_pcs; a.N = p.N();
    a.nzeta = pr->nzeta; a.m = pr->m; a.model = pr->model;
    a.alpha = pr->alpha; a.beta = pr->beta; a.u = pr->u; a.M = pr->M;
    a.w = pr->nw > 0 ? pr->w : nullptr; a.nw = pr->nw;
    return a;
}

double eps_of(double x) {   // MATLAB eps(x): spacing of doubles at |x|
    x = std::fabs(x);
    if (!(x > 0) || !std::isfinite(x)) return std::ldexp(1.0, -1074);
    int e = 0;
    std::frexp(x, &e);      // x = f * 2^e, f in [0.5, 1)
    return std::ldexp(1.0, e - 53);
}

// ---------------------------------------------------------------------------------------------------------------
// Gram-route refinement: one more data pass over a shard (device pointers in `pr`).  The regressors of a chunk are
// materialised feature-major ([Px | Py] rows x Mc snapshots, L2-resident), transformed by the current basis,
//     Z = S Px            (DMMA, B operand k-major: kf_launch_gemm_bkmajor),
// and the Gram / cross products of the NEW features are accumulated:  G2 += Z Z',  C2 += Z Py'.  Z has an almost
// orthonormal leading block and small residual columns behind it, so G2 carries the information that the plain Gram
// G = Px'Px loses to cond(Px)^2 * eps  (Ksysid.m:1069 solves by QR and never squares the condition number).
int refine_pass(kf_ctx* ctx, const kf_problem* pr) {
    KF_TRY(check_problem(ctx, pr));
    KfRefine& R = ctx->rf;
    const KfLayout& L = ctx->lay;
    if (!L.valid || !R.pending || !same_layout(L, ctx->prog, pr, 0)) {
        ctx->err = "refinement pass: no pending refinement for this problem layout";
        return KF_EINVAL;
    }
    return dense_pass(ctx, pr, R.d_St.as<double>(), R.d_G2C2.as<double>(), R.Mc, 0, -1);
}

// Dense accumulation over the chunks [c0, c1) of a shard:  GC = [G | C] (each Pp x Pp) += Z Z' and Z Py', with Z = S Px
// (St = S' column-major, the refinement basis) or Z = Px (St == nullptr: the plain dense Gram of a `loaded` model).
int dense_pass(kf_ctx* ctx, const kf_problem* pr, const double* St, double* GC, int Mc, long long c0, long long c1) {
    KfRefine& R = ctx->rf;
    const KfLayout& L = ctx->lay;
    cudaStream_t st = ctx->stream;
    const int P = L.P, Pp = L.Pp;
    const long long rp_rows = 2LL * P + KF_BM;          // rows [0,P) Px, [P,2P) Py, finite padding behind (k range runs to Pp)
    const bool fresh = R.d_RP.bytes < (size_t)rp_rows * Mc * sizeof(double);
    KF_CUDA(ctx, R.d_RP.ensure((size_t)rp_rows * Mc * sizeof(double)));
    if (St) KF_CUDA(ctx, R.d_Z.ensure((size_t)Pp * Mc * sizeof(double)));
    if (fresh) KF_CUDA(ctx, cudaMemsetAsync(R.d_RP.p, 0, (size_t)rp_rows * Mc * sizeof(double), st));
    if (ctx->prog.n_pcs) KF_CUDA(ctx, ctx->d_full.ensure((size_t)2 * ctx->prog.n_full() * Mc * sizeof(double)));
    double* RP = R.d_RP.as<double>();
    double* Z = St ? R.d_Z.as<double>() : RP;
    double* G2 = GC;
    double* C2 = G2 + (size_t)Pp * Pp;
    const long long nchunks = (pr->M + Mc - 1) / Mc;
    if (c1 < 0 || c1 > nchunks) c1 = nchunks;
    for (long long c = c0; c < c1; ++c) {
        const long long count = std::min<long long>(Mc, pr->M - c * Mc);
        if (count < Mc) KF_CUDA(ctx, cudaMemsetAsync(RP, 0, (size_t)2 * P * Mc * sizeof(double), st));   // tail columns contribute zeros
        KfLiftArgs a = lift_args_of(ctx, pr);
        a.N = L.N;
        a.start = c * Mc;
        a.Mc = Mc;
        a.full = ctx->d_full.as<double>();
        KF_TRY(kf_launch_regressors(ctx, a, RP, nullptr, Mc, st));
        if (St) {
            KfGemmGrid g{};
            g.A = St; g.lda = Pp;                            // A[q][i] = S(q, i): row q of S contiguous in i (= S' column-major)
            g.B = RP; g.ldb = Mc;                            // B[i][snapshot]
            g.out = Z; g.ldm = Mc; g.ldn = 1;
            g.m = P; g.n = Mc; g.k0 = 0; g.k1 = Pp; g.alpha = 1.0; g.accumulate = 0;
            KF_TRY(kf_launch_gemm_bkmajor(ctx, g, st));
        }
        KfGemmGrid gg{};                                 // G2(m, n) += sum_s Z[m][s] Z[n][s], lower tiles
        gg.A = Z; gg.B = Z; gg.lda = gg.ldb = Mc; gg.out = G2; gg.ldm = 1; gg.ldn = Pp;
        gg.m = P; gg.n = P; gg.k0 = 0; gg.k1 = Mc; gg.alpha = 1.0; gg.accumulate = 1; gg.lower_only = 1;
        KF_TRY(kf_launch_gemm_grid(ctx, gg, st));
        KfGemmGrid gc{};                                 // C2(m, n) += sum_s Z[m][s] Py[n][s]
        gc.A = Z; gc.B = RP + (size_t)P * Mc; gc.lda = gc.ldb = Mc; gc.out = C2; gc.ldm = 1; gc.ldn = Pp;
        gc.m = P; gc.n = L.Pc; gc.k0 = 0; gc.k1 = Mc; gc.alpha = 1.0; gc.accumulate = 1;
        KF_TRY(kf_launch_gemm_grid(ctx, gc, st));
    }
    return KF_OK;
}

// start / continue the refinement after a factorisation (W = factor, not cleaned; d_perm its pivot order, r its rank)
int refine_advance(kf_ctx* ctx, int P, int Pp, int r, const int* d_perm) {
    KfRefine& R = ctx->rf;
    cudaStream_t st = ctx->stream;
    const size_t mat = (size_t)Pp * Pp * sizeof(double);
    KF_CUDA(ctx, R.d_S.ensure(mat));
    KF_CUDA(ctx, R.d_St.ensure(mat));
    KF_CUDA(ctx, R.d_Sp.ensure(mat));
    KF_CUDA(ctx, R.d_G2C2.ensure(2 * mat));
    if (R.level == 0) {
        KF_TRY(kf_rf_identity(ctx, R.d_S.as<double>(), P, Pp, st));
        R.orig.resize(P);
        for (int i = 0; i < P; ++i) R.orig[i] = i;
    }
    KF_TRY(kf_rf_update_basis(ctx, P, Pp, ctx->d_W.as<double>(), d_perm, r, R.d_S.as<double>(), R.d_Sp.as<double>(), st));
    KF_TRY(kf_rf_transpose(ctx, R.d_S.as<double>(), R.d_St.as<double>(), Pp, st));
    std::vector<int> perm(P), o2(P);
    KF_CUDA(ctx, cudaMemcpyAsync(perm.data(), d_perm, sizeof(int) * P, cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaMemsetAsync(R.d_G2C2.p, 0, 2 * mat, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    for (int j = 0; j < P; ++j) o2[j] = R.orig[perm[j]];
    R.orig.swap(o2);
    R.forced = r;
    R.pending = true;
    // chunk of a refinement pass: [Px | Py] rows + Z rows of Mc snapshots should stay L2-resident
    long long Mc = (long long)(ctx->opt_panel_mb * 1048576.0 / (8.0 * (2.0 * P + Pp + KF_BM)));
    Mc = std::max<long long>(256, std::min<long long>(Mc / 256 * 256, 8192));
    R.Mc = (int)Mc;
    return KF_OK;
}

// One step of the Gram-route least-squares solve (K = Px \ Py, Ksysid.m:1069) from the reduced matrices on the device.
// Returns KF_OK with K (Pp x Pp, ld Pp, rows in ORIGINAL column order) in *K_out, or KF_EAGAIN when another data pass
// (refine_pass over every shard + all-reduce of rf.d_G2C2) is needed first.
int gram_ls_step(kf_ctx* ctx, const kf_solve* sv, kf_result* out, int* d_perm, double** K_out) {
    const KfLayout& L = ctx->lay;
    KfRefine& R = ctx->rf;
    cudaStream_t st = ctx->stream;
    const int P = L.P, Pp = L.Pp;
    const size_t mat = (size_t)Pp * Pp * sizeof(double);
    const double tol = sv->pivot_tol > 0 ? sv->pivot_tol : 1e-7;
    const double ltol = ctx->opt_refine_level_tol;
    std::vector<double> piv(P);
    int rank = 0;
    double minp = 0, maxp = 0, rej = -1;
    auto ratio = [&](int a, int b) {   // max / min of the accepted pivots [a, b)
        double lo = 0, hi = 0;
        for (int j = a; j < b; ++j) {
            lo = (j == a) ? piv[j] : std::min(lo, piv[j]);
            hi = std::max(hi, piv[j]);
        }
        return (b > a && lo > 0) ? hi / lo : 1.0;
    };
    if (!R.pending) {
        // ---- first factorisation: G itself
        KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_W.p, ctx->d_G.p, mat, cudaMemcpyDeviceToDevice, st));
        KF_TRY(kf_pchol_factor(ctx, P, Pp, ctx->d_W.as<double>(), tol, 0.0, 0, d_perm, &rank, &minp, &maxp, piv.data(), st, &rej));
        const double kappa = ratio(0, rank);
        R.level = 0;
        R.cond_est = kappa;
        R.r11 = maxp;
        R.min_piv = minp;
        out->info.cond_est = kappa;
        out->info.max_pivot = maxp;
        const bool refine = rank > 0 && (ctx->opt_refine == 2 || (ctx->opt_refine == 1 && kappa > ctx->opt_refine_kappa));
        if (!refine) {
            KF_TRY(kf_pchol_solve(ctx, P, Pp, L.Pc, ctx->d_W.as<double>(), ctx->d_C.as<double>(), ctx->d_K.as<double>(), d_perm, rank, 1, st));
            out->info.rank = rank;
            out->info.min_pivot = minp;
            out->info.refine_passes = 0;
            *K_out = ctx->d_K.as<double>();
            return KF_OK;
        }
        // level 0 keeps the pivots within the level's dynamic range of |R_11|; the factor of that prefix is the one just
        // computed unless the user tolerance let smaller pivots in
        if (minp < ltol * maxp) {
            KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_W.p, ctx->d_G.p, mat, cudaMemcpyDeviceToDevice, st));
            KF_TRY(kf_pchol_factor(ctx, P, Pp, ctx->d_W.as<double>(), ltol, 0.0, 0, d_perm, &rank, &minp, &maxp, piv.data(), st, &rej));
            R.min_piv = minp;
        }
        KF_TRY(refine_advance(ctx, P, Pp, rank, d_perm));
        return KF_EAGAIN;
    }
    // ---- a refinement pass has been accumulated (and reduced): factor the Gram of the new features
    double* G2 = R.d_G2C2.as<double>();
    double* C2 = G2 + (size_t)Pp * Pp;
    KF_TRY(kf_rf_symmetrize(ctx, G2, Pp, st));
    KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_W.p, G2, mat, cudaMemcpyDeviceToDevice, st));
    // rank tolerance of mldivide: |R_jj| <= max(size(Px)) * eps(|R_11|); an explicit pivot_tol stays a relative floor
    const double mtol = (double)std::max<long long>(R.M_total, P) * eps_of(R.r11);
    const double tolabs = std::max(mtol, sv->pivot_tol > 0 ? sv->pivot_tol * R.r11 : 0.0);
    KF_TRY(kf_pchol_factor(ctx, P, Pp, ctx->d_W.as<double>(), ltol, tolabs, R.forced, d_perm, &rank, &minp, &maxp, piv.data(), st, &rej));
    R.level += 1;
    const double kp = ratio(0, R.forced), kr = ratio(R.forced, rank);
    for (int j = R.forced; j < rank; ++j) R.min_piv = std::min(R.min_piv, piv[j]);   // residual pivots are true |R_jj|
    // stopped at the level floor with candidates left above the rank tolerance: they need a level of their own
    const bool unresolved = rank < P && rej > tolabs;
    const bool converged = std::max(kp, kr) <= ctx->opt_refine_kappa && !unresolved;
    if (!converged && R.level < ctx->opt_refine_max) {
        KF_TRY(refine_advance(ctx, P, Pp, rank, d_perm));
        return KF_EAGAIN;
    }
    // ---- solve in the new basis and map back:  K = (Pi S)' K_Z
    KF_TRY(kf_rf_gather_rows(ctx, R.d_S.as<double>(), d_perm, P, Pp, R.d_Sp.as<double>(), st));
    KF_TRY(kf_pchol_solve(ctx, P, Pp, L.Pc, ctx->d_W.as<double>(), C2, ctx->d_K.as<double>(), d_perm, rank, 0, st));
    KF_CUDA(ctx, ctx->d_tmp.ensure(mat));
    KF_CUDA(ctx, cudaMemsetAsync(ctx->d_tmp.p, 0, mat, st));
    {
        KfGemmGrid g{};
        g.A = R.d_Sp.as<double>(); g.lda = Pp;           // A[i][j] = Sp(j, i): column i of Sp, j contiguous
        g.B = ctx->d_K.as<double>(); g.ldb = Pp;         // B[c][j] = K_Z(j, c)
        g.out = ctx->d_tmp.as<double>(); g.ldm = 1; g.ldn = Pp;
        g.m = P; g.n = L.Pc; g.k0 = 0; g.k1 = (int)kf_roundup(rank, KF_BK); g.alpha = 1.0; g.accumulate = 0;
        KF_TRY(kf_launch_gemm_grid(ctx, g, st));
    }
    // pivot order in original column indices
    {
        std::vector<int> perm(P), o2(P);
        KF_CUDA(ctx, cudaMemcpyAsync(perm.data(), d_perm, sizeof(int) * P, cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        for (int j = 0; j < P; ++j) o2[j] = R.orig[perm[j]];
        KF_CUDA(ctx, cudaMemcpyAsync(d_perm, o2.data(), sizeof(int) * P, cudaMemcpyHostToDevice, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
    }
    out->info.rank = rank;
    out->info.min_pivot = R.min_piv;
    out->info.max_pivot = R.r11;
    out->info.cond_est = R.cond_est;
    out->info.refine_passes = R.level;
    out->info.refine_capped = converged ? 0 : 1;
    R.pending = false;
    *K_out = ctx->d_tmp.as<double>();
    return KF_OK;
}

// assemble G, C from the (already slab-reduced, possibly all-reduced) accumulator and solve
int solve_from_accum(kf_ctx* ctx, const kf_solve* sv, kf_result* out) {
    const KfLayout& L = ctx->lay;
    if (!L.valid) {
        ctx->err = "kf_solve_dev: nothing accumulated";
        return KF_EINVAL;
    }
    cudaStream_t st = ctx->stream;
    const int P = L.P, Pp = L.Pp;
    const size_t mat = (size_t)Pp * Pp * sizeof(double);
    KF_CUDA(ctx, ctx->d_G.ensure(mat));
    KF_CUDA(ctx, ctx->d_C.ensure(mat));
    KF_CUDA(ctx, ctx->d_K.ensure(mat));
    KF_CUDA(ctx, ctx->d_W.ensure(mat));
    KF_CUDA(ctx, ctx->d_misc.ensure(4096));
    KF_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
    if (!ctx->rf.pending) {
        if (L.oz) {
            KF_TRY(kf_oz_to_gc(ctx, L, ctx->rf.d_dense.as<double>(), ctx->d_G.as<double>(), ctx->d_C.as<double>(), st));
        } else if (L.dense) {
            double* GC = ctx->rf.d_dense.as<double>();
            KF_TRY(kf_rf_symmetrize(ctx, GC, Pp, st));
            KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_G.p, GC, mat, cudaMemcpyDeviceToDevice, st));
            KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_C.p, GC + (size_t)Pp * Pp, mat, cudaMemcpyDeviceToDevice, st));
        } else
        KF_TRY(kf_assemble(ctx, ctx->d_accum.as<double>(), ctx->d_tilemeta.as<KfTile>(), (int)L.tiles.size(), L,
                           ctx->d_G.as<double>(), ctx->d_C.as<double>(), st));
        KF_TRY(copy_out_matrix(ctx, ctx->d_G.as<double>(), Pp, P, out->G));
        KF_TRY(copy_out_matrix(ctx, ctx->d_C.as<double>(), Pp, P, out->C));
        // total snapshot count: the trailer of the accumulator if it went through an all-reduce, else this rank's count
        double tr = 0;
        KF_CUDA(ctx, cudaMemcpyAsync(&tr, accum_base(ctx) + (L.slab - KF_ACC_TRAILER), sizeof(double), cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->rf.M_total = tr >= 1.0 ? (long long)(tr + 0.5) : ctx->accum_M;
    }
    if (!sv) {
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        return KF_OK;
    }
    int* d_perm = reinterpret_cast<int*>(ctx->d_misc.as<char>() + 1024);
    if ((size_t)P * sizeof(int) + 1024 > ctx->d_misc.bytes) {
        KF_CUDA(ctx, ctx->d_misc.ensure(1024 + (size_t)Pp * sizeof(int)));
        d_perm = reinterpret_cast<int*>(ctx->d_misc.as<char>() + 1024);
    }
    if (sv->least_squares) {
        int method = sv->ls_method == KF_LS_AUTO ? KF_LS_GRAM : sv->ls_method;
        if (method != KF_LS_GRAM) {
            ctx->err = "kf_solve_dev: only KF_LS_GRAM can solve from the accumulator (KF_LS_QR needs the snapshots: use kf_fit)";
            return KF_EINVAL;
        }
        double* Kdev = nullptr;
        const int rc = gram_ls_step(ctx, sv, out, d_perm, &Kdev);
        if (rc == KF_EAGAIN) {
            float ms0 = 0.f;
            KF_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
            KF_CUDA(ctx, cudaStreamSynchronize(st));
            KF_CUDA(ctx, cudaEventElapsedTime(&ms0, ctx->ev[4], ctx->ev[5]));
            out->info.t_solve_ms += ms0;
            return KF_EAGAIN;
        }
        if (rc) return rc;
        out->info.ls_method_used = KF_LS_GRAM;
        KF_TRY(copy_out_matrix(ctx, Kdev, Pp, P, out->K, L.Pc));
        if (out->perm) KF_CUDA(ctx, cudaMemcpyAsync(out->perm, d_perm, sizeof(int) * P, cudaMemcpyDeviceToHost, st));
    } else {
        if (sv->nt <= 0 || !sv->t) {
            ctx->err = "kf_solve: QP branch needs nt > 0 budgets";
            return KF_EINVAL;
        }
        if (L.Pc != P) {
            ctx->err = "pc_cols (column-restricted K) is only valid for the least-squares branch: the L1 budget couples all P columns";
            return KF_EINVAL;
        }
        // --- `any(eig(G) < 0)` -> G += 1e-6 I (Ksysid.m:1117-1120).  For a numerically singular G the sign
        // of the smallest computed eigenvalue is rounding noise, so AS_REFERENCE shifts exactly when the
        // pivoted Cholesky finds G rank-deficient (a non-positive / negligible pivot).
        const double tol = sv->pivot_tol > 0 ? sv->pivot_tol : 1e-7;
        int rank = 0;
        double minp = 0, maxp = 0;
        KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_W.p, ctx->d_G.p, mat, cudaMemcpyDeviceToDevice, st));
        KF_TRY(kf_solve_gram_ls(ctx, P, Pp, P, ctx->d_W.as<double>(), ctx->d_C.as<double>(), ctx->d_K.as<double>(), tol, d_perm,
                                &rank, &minp, &maxp, st));
        bool shift = sv->psd_shift == KF_PSD_ALWAYS || (sv->psd_shift == KF_PSD_AS_REFERENCE && rank < P);
        if (shift) {
            KF_TRY(kf_add_diag(ctx, ctx->d_G.as<double>(), Pp, P, 1e-6, st));
            KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_W.p, ctx->d_G.p, mat, cudaMemcpyDeviceToDevice, st));
            KF_TRY(kf_solve_gram_ls(ctx, P, Pp, P, ctx->d_W.as<double>(), ctx->d_C.as<double>(), ctx->d_K.as<double>(), 1e-14,
                                    d_perm, &rank, &minp, &maxp, st));   // unconstrained minimiser of the shifted problem
            if (out->G) KF_TRY(copy_out_matrix(ctx, ctx->d_G.as<double>(), Pp, P, out->G));   // the G the QP used
        }
        out->info.rank = rank;
        out->info.psd_shift_applied = shift ? 1 : 0;
        out->info.min_pivot = minp;
        out->info.max_pivot = maxp;
        // --- pinned delay columns (linear model, nd >= 1; Ksysid.m:1139-1164, indices replicated as written)
        int c0 = 0, c1 = 0;
        double pinned_l1 = 0;
        std::vector<double> target;
        if (sv->delay_constraint && L.model == KF_LINEAR && sv->nd >= 1) {
            const int n = sv->n, nd = sv->nd, m = L.m, N = L.N, Nm = P;
            const int nnd = n * nd, mnd = m * nd;
            target.assign((size_t)Nm * (nnd + mnd), 0.0);
            for (int i = 1; i <= nnd; ++i) target[(size_t)(Nm + 1) * (i - 1)] = 1.0;
            for (int i = 1; i <= m; ++i) target[(size_t)Nm * nnd + N + (size_t)(Nm + 1) * (i - 1)] = 1.0;
            for (int i = 1; i <= m * (nd - 1); ++i) {
                const size_t idx = (size_t)Nm * (nnd + m) + nnd + (size_t)(Nm + 1) * (i - 1);
                if (idx < target.size()) target[idx] = 1.0;
            }
            c0 = n;
            c1 = n * (nd + 1) + mnd;
            for (double v : target) pinned_l1 += std::fabs(v);
        }
        // the unconstrained minimiser with the pinned columns in place: ||.||_1 decides which budgets are active
        if (c1 > c0)
            KF_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_K.as<double>() + (size_t)c0 * Pp, (size_t)Pp * sizeof(double), target.data(),
                                           (size_t)P * sizeof(double), (size_t)P * sizeof(double), c1 - c0, cudaMemcpyHostToDevice, st));
        KfQpResult ls{};
        KF_TRY(kf_qp_evaluate(ctx, P, Pp, ctx->d_G.as<double>(), ctx->d_C.as<double>(), ctx->d_K.as<double>(), c0, c1, 0.0, &ls, st));
        std::vector<int> act;   // budgets whose constraint is active
        for (int it = 0; it < sv->nt; ++it) {
            if (sv->t[it] - pinned_l1 < 0) {
                ctx->err = "L1 budget smaller than the pinned delay entries: the QP is infeasible";
                return KF_ENUMERIC;
            }
            if (ls.l1 <= sv->t[it]) {   // inactive: lam = 0, K = unconstrained minimiser
                if (out->objective) out->objective[it] = ls.objective;
                if (out->l1norm) out->l1norm[it] = ls.l1;
                if (out->qp_iters) out->qp_iters[it] = 0;
                if (out->qp_gap) out->qp_gap[it] = std::max(0.0, ls.grad_inner + (sv->t[it] - pinned_l1) * ls.grad_max);
                if (out->K) KF_TRY(copy_out_matrix(ctx, ctx->d_K.as<double>(), Pp, P, out->K + (size_t)it * P * P));
            } else {
                act.push_back(it);
            }
        }
        // all active budgets of the lasso vector are solved in lockstep, in groups that bound the device memory
        // (K_b and dK_b/dlam per budget)
        int capped = 0;
        const size_t per = 2 * mat;
        const int gmax = (int)std::max<size_t>(1, std::min<size_t>(256, (size_t)(12.0 * 1073741824.0) / per));
        for (size_t g0 = 0; g0 < act.size(); g0 += gmax) {
            const int nb = (int)std::min<size_t>(gmax, act.size() - g0);
            KF_CUDA(ctx, ctx->d_Kt.ensure((size_t)nb * mat));
            double* Kt = ctx->d_Kt.as<double>();
            KF_CUDA(ctx, cudaMemsetAsync(Kt, 0, (size_t)nb * mat, st));
            std::vector<double> tf(nb);
            for (int b = 0; b < nb; ++b) {
                tf[b] = sv->t[act[g0 + b]] - pinned_l1;
                if (c1 > c0)
                    KF_CUDA(ctx, cudaMemcpy2DAsync(Kt + (size_t)b * Pp * Pp + (size_t)c0 * Pp, (size_t)Pp * sizeof(double), target.data(),
                                                   (size_t)P * sizeof(double), (size_t)P * sizeof(double), c1 - c0, cudaMemcpyHostToDevice, st));
            }
            std::vector<KfQpResult> qr(nb);
            // auto = the exact active-set solver at every size: on the reference's own small dictionaries the lockstep coordinate
            // descent needs 1e4-1e5 sweeps per multiplier (config 3b, P = 16: 2.0 s against 0.09 s; gaussian 30, P = 68: minutes)
            const bool active_set = ctx->opt_qp_method == 2 || ctx->opt_qp_method == 0;
            // "lasso sweeps are split across GPUs" (north_star): with the library's own communicator the exact active-set sweep is
            // split by COLUMNS of K automatically — rank r factors the columns shard_bounds(P, r, nranks) for all budgets, the
            // step scalars are summed in stream order over NCCL, and the column blocks are exchanged at the end, so every rank
            // ends up with the complete K of every budget.  (A caller-set partition, kf_set_qp_partition, takes precedence.)
            const bool auto_split = active_set && ctx->comm && ctx->nranks > 1 && !(ctx->qp_hi > ctx->qp_lo) && c1 <= c0 && P > 256 &&
                                    P <= 4096 && ctx->opt_qp_split;
            std::vector<size_t> goff, gcnt;
            if (auto_split) {
                const int base = P / ctx->nranks, extra = P % ctx->nranks;
                for (int r = 0; r < ctx->nranks; ++r) {
                    const int lo = r * base + std::min(r, extra), hi = lo + base + (r < extra ? 1 : 0);
                    goff.push_back((size_t)lo * Pp);
                    gcnt.push_back((size_t)(hi - lo) * Pp);
                    if (r == ctx->rank) { ctx->qp_lo = lo; ctx->qp_hi = hi; }
                }
            }
            const bool split = active_set && !auto_split && ctx->qp_hi > ctx->qp_lo;     // caller's column partition across ranks
            double td0 = now_ms();
            if (active_set) {
                // the columns are DEALT round-robin to the ranks (rank r: columns r, r + R, ...; as a contiguous block of a column-
                // permuted C): the support sizes differ systematically between column ranges (config 3a: the psi and the u psi
                // halves), so contiguous blocks of the original order leave ranks idle
                const double* Cs = ctx->d_C.as<double>();
                if (auto_split) {
                    KF_CUDA(ctx, ctx->d_deal.ensure(mat));
                    KF_TRY(kf_qp_deal_cols(ctx, ctx->d_C.as<double>(), ctx->d_deal.as<double>(), P, Pp, ctx->nranks, 1, st));
                    Cs = ctx->d_deal.as<double>();
                }
                td0 = now_ms();
                const int rc_as = kf_solve_l1ball_as(ctx, P, Pp, ctx->d_G.as<double>(), Cs, nb, tf.data(), c0, c1,
                                                     sv->qp_max_iter, Kt, qr.data(), st);
                if (ctx->opt_as_diag) { cudaStreamSynchronize(st); fprintf(stderr, "[as_diag] rank %d: active-set sweep %.1f ms\n", ctx->rank, now_ms() - td0); }
                if (auto_split) { ctx->qp_lo = 0; ctx->qp_hi = 0; }
                if (rc_as) return rc_as;
                if (auto_split)
                    for (int b = 0; b < nb; ++b) {
                        double* Kb = Kt + (size_t)b * Pp * Pp;
                        KF_TRY(kf_comm_gather_blocks(ctx, Kb, goff.data(), gcnt.data(), st));
                        KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_deal.as<double>(), Kb, mat, cudaMemcpyDeviceToDevice, st));   // back to the original column order
                        KF_TRY(kf_qp_deal_cols(ctx, ctx->d_deal.as<double>(), Kb, P, Pp, ctx->nranks, 0, st));
                    }
            }
            else
                KF_TRY(kf_solve_l1ball_multi(ctx, P, Pp, ctx->d_G.as<double>(), ctx->d_C.as<double>(), nb, tf.data(), c0, c1, sv->qp_max_iter,
                                             sv->qp_tol, Kt, qr.data(), st));
            // objective, ||K||_1 and the gap: over ALL columns (the pinned ones count towards the budget row, Ksysid.m:1136);
            // with a column partition the per-rank pieces are combined through the caller's all-reduce
            auto evaluate = [&](int b, KfQpResult* ev) -> int {
                KF_TRY(kf_qp_evaluate(ctx, P, Pp, ctx->d_G.as<double>(), ctx->d_C.as<double>(), Kt + (size_t)b * Pp * Pp, c0, c1, tf[b], ev, st,
                                      split ? ctx->qp_lo : 0, split ? ctx->qp_hi : 0));
                if (split && ctx->qp_allreduce) {
                    double v[3] = {ev->objective, ev->l1, ev->grad_inner}, mx = ev->grad_max;
                    if (ctx->qp_allreduce(ctx->qp_user, v, 3, 0) || ctx->qp_allreduce(ctx->qp_user, &mx, 1, 1)) {
                        ctx->err = "kf_solve: the all-reduce hook failed";
                        return KF_EINVAL;
                    }
                    ev->objective = v[0]; ev->l1 = v[1]; ev->grad_inner = v[2]; ev->grad_max = mx;
                    ev->gap = std::max(0.0, v[2] + std::max(tf[b], 0.0) * mx);
                }
                return KF_OK;
            };
            // Safety net of the auto choice on small problems: a budget the active-set iteration did not settle (step cap, or a
            // certified gap above the 1e-8 objective tolerance — numerically singular Grams with psd_shift = never) sends the group
            // through the coordinate descent, slow but monotone.
            if (active_set && ctx->opt_qp_method == 0 && P <= 256 && !split && !auto_split) {
                bool redo = false;
                for (int b = 0; b < nb && !redo; ++b) {
                    KfQpResult ev{};
                    KF_TRY(evaluate(b, &ev));
                    redo = qr[b].capped || !(ev.gap <= 1e-8 * std::max(std::fabs(ev.objective), 1e-300));
                }
                if (redo)
                    KF_TRY(kf_solve_l1ball_multi(ctx, P, Pp, ctx->d_G.as<double>(), ctx->d_C.as<double>(), nb, tf.data(), c0, c1, sv->qp_max_iter,
                                                 sv->qp_tol, Kt, qr.data(), st));
            }
            const double td1 = now_ms();
            if (ctx->opt_as_diag) { cudaStreamSynchronize(st); fprintf(stderr, "[as_diag] rank %d: sweep + column exchange %.1f ms\n", ctx->rank, td1 - td0); }
            for (int b = 0; b < nb; ++b) {
                const int it = act[g0 + b];
                capped += qr[b].capped;
                KfQpResult ev{};
                KF_TRY(evaluate(b, &ev));
                if (ev.l1 > sv->t[it] * (1.0 + 1e-13) && ev.l1 > pinned_l1) {   // an unconverged iterate: make it feasible
                    KF_TRY(kf_qp_scale_free(ctx, Kt + (size_t)b * Pp * Pp, P, Pp, c0, c1, tf[b] / (ev.l1 - pinned_l1), st));
                    KF_TRY(evaluate(b, &ev));
                }
                if (out->objective) out->objective[it] = ev.objective;
                if (out->l1norm) out->l1norm[it] = ev.l1;
                if (out->qp_iters) out->qp_iters[it] = qr[b].iters;
                if (out->qp_gap) out->qp_gap[it] = ev.gap;
                if (out->K) KF_TRY(copy_out_matrix(ctx, Kt + (size_t)b * Pp * Pp, Pp, P, out->K + (size_t)it * P * P));
            }
            KF_CUDA(ctx, cudaStreamSynchronize(st));
            if (ctx->opt_as_diag) fprintf(stderr, "[as_diag] rank %d: evaluation + copy-out of %d budgets %.1f ms\n", ctx->rank, nb, now_ms() - td1);
        }
        out->info.passes = 1;
        out->info.qp_capped = capped;
        if (out->perm) KF_CUDA(ctx, cudaMemcpyAsync(out->perm, d_perm, sizeof(int) * P, cudaMemcpyDeviceToHost, st));
    }
    KF_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    KF_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    out->info.t_solve_ms += ms;
    ctx->last_solve_ms = (float)out->info.t_solve_ms;
    return KF_OK;
}

}  // namespace

// the fit from DEVICE snapshot pairs (prob->alpha/beta/u are device pointers, column-major, ld = M); the dictionary is prepared.
// accumulated: the caller has already run the lift + Gram of the shard (kf_fit_host_shard overlaps it with the copies).
// host_row0 / host_ld: where this shard's rows sit in the caller's Px / Py (kf_fit_multi: one row block per device).
int kf_fit_device_pairs(kf_ctx* ctx, const kf_problem* prob, const kf_solve* solve, kf_result* out, double t0, bool accumulated,
                        long long host_row0, long long host_ld) {
    const long long M = prob->M;
    const double* d_alpha = prob->alpha;
    const double* d_beta = prob->beta;
    const double* d_u = prob->u;
    if (host_ld <= 0) host_ld = M;
    const bool multi = ctx->comm && ctx->nranks > 1;
    if (!accumulated) {
        ctx->lay.valid = false;
        KF_TRY(accumulate_dev(ctx, prob, true));
    }
    KF_TRY(finish_accum(ctx));
    if (multi) {
        // the ONE data-path collective of a snapshot-sharded fit: sum of the packed partial Grams (+ the snapshot count)
        KF_TRY(write_trailer(ctx));
        KF_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
        KF_TRY(kf_comm_allreduce(ctx, accum_base(ctx), (size_t)ctx->lay.slab, 0, ctx->stream));
        KF_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    }
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    KF_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->last_lift_gram_ms = ms;
    out->info.t_lift_gram_ms = ms;
    out->info.passes = 1;
    ctx->last_allreduce_ms = 0.f;
    if (multi) KF_CUDA(ctx, cudaEventElapsedTime(&ctx->last_allreduce_ms, ctx->ev[2], ctx->ev[3]));
    // optional materialised regressors (koopData.Px / Py, Ksysid.m:1085-1086)
    if (out->Px || out->Py) {
        const int P = ctx->lay.P;
        const long long ldq = kf_qr_ld(M);       // rows padded with zeros to a multiple of 64 (blocked QR, aligned operand rows)
        KF_CUDA(ctx, ctx->d_qr.ensure((size_t)ldq * 2 * P * sizeof(double)));
        KfLiftArgs a = lift_args_of(ctx, prob);
        a.N = ctx->lay.N;
        if (ctx->prog.n_pcs) KF_CUDA(ctx, ctx->d_full.ensure((size_t)2 * a.n_full * M * sizeof(double)));
        a.full = ctx->d_full.as<double>();
        if (ldq > M) KF_CUDA(ctx, cudaMemset2DAsync(ctx->d_qr.as<double>() + M, (size_t)ldq * sizeof(double), 0, (size_t)(ldq - M) * sizeof(double), 2 * (size_t)P, ctx->stream));
        KF_TRY(kf_launch_regressors(ctx, a, ctx->d_qr.as<double>(), nullptr, ldq, ctx->stream));
        const size_t w = (size_t)M * sizeof(double), wq = (size_t)ldq * sizeof(double);
        if (out->Px) KF_CUDA(ctx, cudaMemcpy2DAsync(out->Px + host_row0, (size_t)host_ld * sizeof(double), ctx->d_qr.p, wq, w, P,
                                                    cudaMemcpyDeviceToHost, ctx->stream));
        if (out->Py) KF_CUDA(ctx, cudaMemcpy2DAsync(out->Py + host_row0, (size_t)host_ld * sizeof(double), ctx->d_qr.as<double>() + (size_t)ldq * P,
                                                    wq, w, P, cudaMemcpyDeviceToHost, ctx->stream));
    }
    int method = solve->ls_method;
    if (solve->least_squares && method == KF_LS_AUTO) {
        const double need = (double)M * 2.0 * ctx->lay.P * sizeof(double);
        method = (need <= ctx->opt_qr_max_gb * 1073741824.0) ? KF_LS_QR : KF_LS_GRAM;
    }
    if (multi && method == KF_LS_QR) {
        if (solve->ls_method == KF_LS_QR) {
            ctx->err = "KF_LS_QR factors the materialised regressors of ONE device; a snapshot-sharded fit solves by the (refined) Gram route";
            return KF_EUNSUPPORTED;
        }
        method = KF_LS_GRAM;
    }
    int rc;
    if (solve->least_squares && method == KF_LS_QR) {
        // Householder QRCP on the materialised regressors: the reference's own algorithm (Ksysid.m:1069)
        rc = solve_from_accum(ctx, nullptr, out);      // G, C outputs only
        if (rc) return rc;
        const int P = ctx->lay.P, Pp = ctx->lay.Pp;
        cudaStream_t st = ctx->stream;
        KF_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
        const long long ldq = kf_qr_ld(M);
        KF_CUDA(ctx, ctx->d_qr.ensure((size_t)ldq * 2 * P * sizeof(double)));
        if (!(out->Px || out->Py)) {   // not materialised above
            KfLiftArgs a = lift_args_of(ctx, prob);
            a.N = ctx->lay.N;
            if (ctx->prog.n_pcs) KF_CUDA(ctx, ctx->d_full.ensure((size_t)2 * a.n_full * M * sizeof(double)));
            a.full = ctx->d_full.as<double>();
            if (ldq > M) KF_CUDA(ctx, cudaMemset2DAsync(ctx->d_qr.as<double>() + M, (size_t)ldq * sizeof(double), 0, (size_t)(ldq - M) * sizeof(double), 2 * (size_t)P, st));
            KF_TRY(kf_launch_regressors(ctx, a, ctx->d_qr.as<double>(), nullptr, ldq, st));
        }
        KF_CUDA(ctx, ctx->d_K.ensure((size_t)Pp * Pp * sizeof(double)));
        KF_CUDA(ctx, ctx->d_misc.ensure(1024 + (size_t)Pp * sizeof(int)));
        int* d_perm = reinterpret_cast<int*>(ctx->d_misc.as<char>() + 1024);
        int rank = 0;
        double minp = 0, maxp = 0;
        const int Pc = ctx->lay.Pc;
        KF_TRY(kf_solve_qr_ls(ctx, M, P, Pc, ctx->d_qr.as<double>(), ldq, ctx->d_K.as<double>(), Pp, d_perm, &rank, &minp, &maxp, st));
        out->info.rank = rank;
        out->info.ls_method_used = KF_LS_QR;
        out->info.min_pivot = minp;
        out->info.max_pivot = maxp;
        out->info.passes = 2;
        KF_TRY(copy_out_matrix(ctx, ctx->d_K.as<double>(), Pp, P, out->K, Pc));
        if (out->perm) KF_CUDA(ctx, cudaMemcpyAsync(out->perm, d_perm, sizeof(int) * P, cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        float sms = 0.f;
        KF_CUDA(ctx, cudaEventElapsedTime(&sms, ctx->ev[4], ctx->ev[5]));
        ctx->last_solve_ms = sms;
        out->info.t_solve_ms = sms;
    } else {
        kf_solve sv2 = *solve;
        if (sv2.least_squares) sv2.ls_method = KF_LS_GRAM;
        rc = solve_from_accum(ctx, &sv2, out);
        // ill-conditioned regressor: extra data passes of the multi-level Cholesky-QR refinement (see gram_ls_step)
        ctx->last_refine_ms = 0.f;
        while (rc == KF_EAGAIN) {
            KF_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
            rc = refine_pass(ctx, prob);
            if (rc) break;
            if (multi) {
                rc = kf_comm_allreduce(ctx, ctx->rf.d_G2C2.as<double>(), (size_t)2 * ctx->lay.Pp * ctx->lay.Pp, 0, ctx->stream);
                if (rc) break;
            }
            KF_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
            KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            float rms = 0.f;
            KF_CUDA(ctx, cudaEventElapsedTime(&rms, ctx->ev[2], ctx->ev[3]));
            ctx->last_refine_ms += rms;
            out->info.t_lift_gram_ms += rms;
            rc = solve_from_accum(ctx, &sv2, out);
        }
        out->info.passes = 1 + out->info.refine_passes;
    }
    ctx->lay.valid = false;
    ctx->rf.pending = false;
    out->info.t_total_ms = now_ms() - t0;
    return rc;
}

// The fit of rows [lo, hi) of the HOST snapshot pairs in `prob` (column-major, leading dimension prob->M) on this context's
// device.  The shard is copied in blocks on a copy stream and every block is lifted and contracted as soon as it has arrived,
// so the PCIe transfer hides under the DMMA work of the previous block (pinned host memory makes the copies truly asynchronous;
// pageable memory — what MATLAB hands over — is staged by the driver and still overlaps with the kernels already queued).
int kf_fit_host_shard(kf_ctx* ctx, const kf_basis* basis, const kf_problem* prob, long long lo, long long hi, const kf_solve* solve,
                      kf_result* out) {
    if (!ctx || !out || !solve || !prob) return KF_EINVAL;
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    const double t0 = now_ms();
    std::memset(&out->info, 0, sizeof(out->info));
    KF_TRY(prepare_program(ctx, basis));
    KF_TRY(check_problem(ctx, prob));
    if (lo < 0 || hi > prob->M || hi <= lo) {
        ctx->err = "kf_fit: empty or out-of-range snapshot shard";
        return KF_EINVAL;
    }
    // host -> device: alpha | beta | u, column-major, ld = Mr  (Ksysid.m:1005)
    const long long Mh = prob->M, Mr = hi - lo;
    const size_t nz = (size_t)Mr * prob->nzeta, nu = (size_t)Mr * prob->m, nwd = (size_t)Mr * prob->nw;
    KF_CUDA(ctx, ctx->d_in.ensure((2 * nz + nu + nwd + 2) * sizeof(double)));
    double* d_alpha = ctx->d_in.as<double>();
    double* d_beta = d_alpha + nz;
    double* d_u = d_beta + nz;
    double* d_w = d_u + nu;
    kf_problem dp = *prob;
    dp.M = Mr;
    dp.alpha = d_alpha;
    dp.beta = d_beta;
    dp.u = d_u;
    dp.w = prob->nw > 0 ? d_w : nullptr;
    // layout first: the copy blocks are whole chunks of the lifted panel
    ctx->lay.valid = false;
    KF_TRY(accumulate_dev(ctx, &dp, true, 0, 0));
    const long long Mc = ctx->lay.Mc, nchunks = (Mr + Mc - 1) / Mc;
    long long per = std::max<long long>(2, (nchunks + 15) / 16);       // <= 16 blocks, an even number of chunks each
    per += per & 1;
    const long long nblk = (nchunks + per - 1) / per;
    while ((long long)ctx->copy_ev.size() < nblk) {
        cudaEvent_t e = nullptr;
        KF_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->copy_ev.push_back(e);
    }
    // the copies may only overwrite d_in once everything queued on the context stream (an earlier fit) is done
    KF_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    KF_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork, 0));
    auto copy_block = [&](long long b) -> int {
        const long long r0 = b * per * Mc, r1 = std::min(Mr, (b + 1) * per * Mc);
        const size_t w = (size_t)(r1 - r0) * sizeof(double);
        cudaStream_t cs = ctx->copy_stream;
        KF_CUDA(ctx, cudaMemcpy2DAsync(d_alpha + r0, (size_t)Mr * sizeof(double), prob->alpha + lo + r0, (size_t)Mh * sizeof(double), w,
                                       prob->nzeta, cudaMemcpyHostToDevice, cs));
        KF_CUDA(ctx, cudaMemcpy2DAsync(d_beta + r0, (size_t)Mr * sizeof(double), prob->beta + lo + r0, (size_t)Mh * sizeof(double), w,
                                       prob->nzeta, cudaMemcpyHostToDevice, cs));
        if (prob->m > 0)
            KF_CUDA(ctx, cudaMemcpy2DAsync(d_u + r0, (size_t)Mr * sizeof(double), prob->u + lo + r0, (size_t)Mh * sizeof(double), w, prob->m,
                                           cudaMemcpyHostToDevice, cs));
        if (prob->nw > 0)
            KF_CUDA(ctx, cudaMemcpy2DAsync(d_w + r0, (size_t)Mr * sizeof(double), prob->w + lo + r0, (size_t)Mh * sizeof(double), w, prob->nw,
                                           cudaMemcpyHostToDevice, cs));
        KF_CUDA(ctx, cudaEventRecord(ctx->copy_ev[b], cs));
        return KF_OK;
    };
    KF_TRY(copy_block(0));
    for (long long b = 0; b < nblk; ++b) {
        if (b + 1 < nblk) KF_TRY(copy_block(b + 1));
        KF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[b], 0));
        KF_TRY(accumulate_dev(ctx, &dp, false, b * per, std::min(nchunks, (b + 1) * per)));
    }
    return kf_fit_device_pairs(ctx, &dp, solve, out, t0, true, lo, Mh);
}

// ====================================================================== C ABI
extern "C" {

int kf_version(void) { return KF_VERSION; }

const char* kf_last_error(const kf_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int kf_create(kf_ctx** out, int device) {
    if (!out) return KF_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_err = std::string("kf_create: no CUDA device (") + cudaGetErrorString(e) + "); koopfit has no CPU fallback";
        return KF_ECUDA;
    }
    if (device < 0 || device >= count) {
        g_create_err = "kf_create: device index out of range";
        return KF_EINVAL;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return KF_ECUDA;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10) {
        g_create_err = "kf_create: libkoopfit.so is built for sm_100a (B200) only; found compute capability " +
                       std::to_string(prop.major) + "." + std::to_string(prop.minor);
        return KF_EUNSUPPORTED;
    }
    kf_ctx* ctx = new kf_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->stream_x[0], cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->stream_x[1], cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 9 && ok; ++i) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
    for (int i = 0; i < KF_MAX_PIPES && ok; ++i) ok = cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        g_create_err = "kf_create: stream/event creation failed";
        delete ctx;
        return KF_ECUDA;
    }
    *out = ctx;
    return KF_OK;
}

void kf_destroy(kf_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    // this context's streams only: a device-wide synchronisation would invalidate a CUDA-graph capture of another context
    for (cudaStream_t q : {ctx->stream, ctx->stream2, ctx->stream_x[0], ctx->stream_x[1], ctx->copy_stream})
        if (q) cudaStreamSynchronize(q);
    KfBuf* bufs[] = {&ctx->d_order, &ctx->d_ops, &ctx->d_centres, &ctx->d_pcs, &ctx->d_panel[0], &ctx->d_panel[1], &ctx->d_panel[2], &ctx->d_panel[3], &ctx->d_full,
                     &ctx->d_tasks[0], &ctx->d_tasks[1], &ctx->d_tma_tasks[0], &ctx->d_tma_tasks[1], &ctx->d_accum, &ctx->d_tilemeta, &ctx->d_G, &ctx->d_C, &ctx->d_K,
                     &ctx->d_W, &ctx->d_in, &ctx->d_misc, &ctx->d_qr, &ctx->d_tmp, &ctx->d_K2, &ctx->d_K3, &ctx->d_Kt,
                     &ctx->d_as_mat, &ctx->d_as_aux, &ctx->d_as_ws, &ctx->d_series, &ctx->d_lift_groups, &ctx->d_bqr, &ctx->d_deal, &ctx->d_pca,
                     &ctx->rf.d_S, &ctx->rf.d_St, &ctx->rf.d_Sp, &ctx->rf.d_G2C2, &ctx->rf.d_RP, &ctx->rf.d_Z, &ctx->rf.d_dense,
                     &ctx->rf.d_dense2, &ctx->rf.d_dense_x[0], &ctx->rf.d_dense_x[1]};
    kf_oz_destroy(ctx);
    if (ctx->pchol_graph.exec) cudaGraphExecDestroy(ctx->pchol_graph.exec);
    for (KfBuf* b : bufs) b->release();
    for (int i = 0; i < 9; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    kf_comm_destroy(ctx);
    for (cudaEvent_t e : ctx->copy_ev) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    for (int i = 0; i < 2; ++i)
        if (ctx->stream_x[i]) cudaStreamDestroy(ctx->stream_x[i]);
    for (int i = 0; i < KF_MAX_PIPES; ++i)
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int kf_basis_dims(const kf_basis* basis, int model, int m, int* n_full, int* N, int* P) {
    KfProgram prog;
    std::string err;
    int rc = kf_build_program(basis, prog, err);
    if (rc) {
        g_create_err = err;
        return rc;
    }
    if (n_full) *n_full = prog.n_full();
    if (N) *N = prog.N();
    if (P) *P = kf_regressor_width(model, prog.N(), m);
    return KF_OK;
}

int kf_block_table(int type, int degree, int nv, int* rows, int* cols, int* table) {
    std::vector<int> t;
    std::string err;
    int r = 0, c = 0;
    int rc = kf_block_rows(type, degree, nv, &r, &c, table ? &t : nullptr, err);
    if (rc) {
        g_create_err = err;
        return rc;
    }
    if (rows) *rows = r;
    if (cols) *cols = c;
    if (table && !t.empty()) std::memcpy(table, t.data(), sizeof(int) * t.size());
    return KF_OK;
}

int kf_lift(kf_ctx* ctx, const kf_basis* basis, long long rows, const double* V, double* Psi) {
    if (!ctx) return KF_EINVAL;
    if (rows <= 0 || !V || !Psi) {
        ctx->err = "kf_lift: rows > 0, V and Psi required";
        return KF_EINVAL;
    }
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    KF_TRY(prepare_program(ctx, basis));
    const KfProgram& p = ctx->prog;
    const size_t vin = (size_t)rows * p.nv * sizeof(double), vout = (size_t)rows * p.N() * sizeof(double);
    const size_t vfull = p.n_pcs ? (size_t)rows * p.n_full() * sizeof(double) : 0;
    KF_CUDA(ctx, ctx->d_in.ensure(vin));
    KF_CUDA(ctx, ctx->d_tmp.ensure(vout));
    if (vfull) KF_CUDA(ctx, ctx->d_full.ensure(vfull));
    KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_in.p, V, vin, cudaMemcpyHostToDevice, ctx->stream));
    KF_TRY(kf_launch_lift_points(ctx, ctx->d_ops.as<KfOp>(), ctx->d_centres.as<double>(), ctx->d_pcs.as<double>(), p.nv,
                                 p.n_full(), p.n_pcs, ctx->d_in.as<double>(), rows, ctx->d_full.as<double>(),
                                 ctx->d_tmp.as<double>(), rows, ctx->stream));
    KF_CUDA(ctx, cudaMemcpyAsync(Psi, ctx->d_tmp.p, vout, cudaMemcpyDeviceToHost, ctx->stream));
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KF_OK;
}

int kf_pca(kf_ctx* ctx, const kf_basis* basis, long long rows, const double* V, double* mu, double* latent, double* coeff) {
    if (!ctx) return KF_EINVAL;
    if (rows < 2 || !V) {
        ctx->err = "kf_pca: rows >= 2 and V required";
        return KF_EINVAL;
    }
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    KF_TRY(prepare_program(ctx, basis));
    const size_t vin = (size_t)rows * ctx->prog.nv * sizeof(double);
    KF_CUDA(ctx, ctx->d_in.ensure(vin));
    KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_in.p, V, vin, cudaMemcpyHostToDevice, ctx->stream));
    return kf_pca_points(ctx, rows, ctx->d_in.as<double>(), mu, latent, coeff);
}

int kf_accumulate_dev(kf_ctx* ctx, const kf_basis* basis, const kf_problem* prob, int reset) {
    if (!ctx) return KF_EINVAL;
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!reset && ctx->lay.valid && ctx->rf.pending) return refine_pass(ctx, prob);   // kf_solve_dev returned KF_EAGAIN
    if (reset || !ctx->lay.valid) KF_TRY(prepare_program(ctx, basis));
    return accumulate_dev(ctx, prob, reset != 0 || !ctx->lay.valid);
}

int kf_regressors_dev(kf_ctx* ctx, const kf_basis* basis, const kf_problem* prob, double* dev_PxPy, long long ld) {
    if (!ctx || !prob || !dev_PxPy) return KF_EINVAL;
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    KF_TRY(prepare_program(ctx, basis));
    ctx->lay.valid = false;
    KF_TRY(check_problem(ctx, prob));
    if (ld < prob->M) { ctx->err = "kf_regressors_dev: ld < M"; return KF_EINVAL; }
    KfLiftArgs a = lift_args_of(ctx, prob);
    if (ctx->prog.n_pcs) KF_CUDA(ctx, ctx->d_full.ensure((size_t)2 * a.n_full * prob->M * sizeof(double)));
    a.full = ctx->d_full.as<double>();
    return kf_launch_regressors(ctx, a, dev_PxPy, nullptr, ld, ctx->stream);
}

int kf_lift_dev(kf_ctx* ctx, const kf_basis* basis, long long rows, const double* dev_V, double* dev_Psi) {
    if (!ctx || rows < 0 || (rows && (!dev_V || !dev_Psi))) return KF_EINVAL;
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    KF_TRY(prepare_program(ctx, basis));
    ctx->lay.valid = false;
    const KfProgram& p = ctx->prog;
    if (p.n_pcs) KF_CUDA(ctx, ctx->d_full.ensure((size_t)p.n_full() * rows * sizeof(double)));
    return kf_launch_lift_points(ctx, ctx->d_ops.as<KfOp>(), ctx->d_centres.as<double>(), ctx->d_pcs.as<double>(), p.nv, p.n_full(),
                                 p.n_pcs, dev_V, rows, ctx->d_full.as<double>(), dev_Psi, rows, ctx->stream);
}

int kf_accum_buffer(kf_ctx* ctx, double** dev_ptr, size_t* count) {
    if (!ctx || !ctx->lay.valid) return KF_EINVAL;
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->rf.pending) {         // a refinement pass: the Gram / cross products of the new features
        if (dev_ptr) *dev_ptr = ctx->rf.d_G2C2.as<double>();
        if (count) *count = (size_t)2 * ctx->lay.Pp * ctx->lay.Pp;
        return KF_OK;
    }
    KF_TRY(finish_accum(ctx));     // all slabs -> slab 0 (the folded slabs are zeroed, so this is idempotent)
    KF_TRY(write_trailer(ctx));    // + this rank's snapshot count, summed by the same all-reduce
    if (dev_ptr) *dev_ptr = accum_base(ctx);
    if (count) *count = (size_t)ctx->lay.slab;
    return KF_OK;
}

int kf_solve_dev(kf_ctx* ctx, const kf_solve* solve, kf_result* out) {
    if (!ctx || !out) return KF_EINVAL;
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    const double t0 = now_ms();
    if (!ctx->rf.pending) {
        std::memset(&out->info, 0, sizeof(out->info));
        KF_TRY(finish_accum(ctx));
    }
    int rc = solve_from_accum(ctx, solve, out);
    if (rc == KF_EAGAIN) return rc;   // another data pass: kf_accumulate_dev(reset = 0), all-reduce kf_accum_buffer, call again
    out->info.passes = 1 + out->info.refine_passes;
    ctx->lay.valid = false;        // accumulator consumed: the next accumulate must reset
    ctx->rf.pending = false;
    out->info.t_total_ms = now_ms() - t0;
    return rc;
}

int kf_sync(kf_ctx* ctx) {
    if (!ctx) return KF_EINVAL;
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream2));
    for (int i = 0; i < 2; ++i) KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream_x[i]));
    return KF_OK;
}

void* kf_stream(kf_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int kf_fit(kf_ctx* ctx, const kf_basis* basis, const kf_problem* prob, const kf_solve* solve, kf_result* out) {
    if (!ctx || !out || !solve || !prob) return KF_EINVAL;
    return kf_fit_host_shard(ctx, basis, prob, 0, prob->M, solve, out);
}

int kf_fit_dev(kf_ctx* ctx, const kf_basis* basis, const kf_problem* prob, const kf_solve* solve, kf_result* out) {
    if (!ctx || !out || !solve || !prob) return KF_EINVAL;
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    const double t0 = now_ms();
    std::memset(&out->info, 0, sizeof(out->info));
    KF_TRY(prepare_program(ctx, basis));
    KF_TRY(check_problem(ctx, prob));
    return kf_fit_device_pairs(ctx, prob, solve, out, t0, false, 0, 0);
}

long long kf_series_pairs(long long T, int nd, const double* t) {
    if (!t || T - nd - 1 < 1) return 0;
    long long kept = 0;
    for (long long j = 0; j < T - nd - 1; ++j) kept += t[nd + j] < t[nd + j + 1] ? 1 : 0;
    return kept > 0 ? kept - 1 : 0;
}

int kf_fit_series(kf_ctx* ctx, const kf_basis* basis, const kf_series* ser, const kf_solve* solve, kf_scale* scale, kf_result* out) {
    if (!ctx || !out || !solve) return KF_EINVAL;
    if (!ser || ser->T < 3 || ser->n < 1 || ser->m < 0 || ser->nd < 0 || !ser->t || !ser->y || (ser->m > 0 && !ser->u)) {
        ctx->err = "kf_fit_series: T >= 3, n >= 1, t, y (and u when m > 0) required";
        return KF_EINVAL;
    }
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    const double t0 = now_ms();
    std::memset(&out->info, 0, sizeof(out->info));
    KF_TRY(prepare_program(ctx, basis));
    cudaStream_t st = ctx->stream;
    const long long T = ser->T;
    const int n = ser->n, m = ser->m, nd = ser->nd, ncol = n + m;
    const int nzeta = n * (nd + 1) + m * nd;          // Ksysid.m:86
    // device: t | y | u | scale
    KF_CUDA(ctx, ctx->d_series.ensure(((size_t)T * (1 + ncol) + 2 * ncol + 2) * sizeof(double)));
    double* d_t = ctx->d_series.as<double>();
    double* d_y = d_t + T;
    double* d_u = d_y + (size_t)T * n;
    double* d_scale = d_u + (size_t)T * m;
    KF_CUDA(ctx, cudaMemcpyAsync(d_t, ser->t, (size_t)T * sizeof(double), cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(d_y, ser->y, (size_t)T * n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (m) KF_CUDA(ctx, cudaMemcpyAsync(d_u, ser->u, (size_t)T * m * sizeof(double), cudaMemcpyHostToDevice, st));
    if (ser->prescaled) {
        std::vector<double> one(2 * ncol, 0.0);
        for (int c = 0; c < ncol; ++c) one[ncol + c] = 1.0;
        KF_CUDA(ctx, cudaMemcpyAsync(d_scale, one.data(), one.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
    } else {
        KF_TRY(kf_pp_scale(ctx, d_y, d_u, T, n, m, d_scale, st));
    }
    long long M = 0;
    KF_TRY(kf_pp_count_pairs(ctx, d_t, T, nd, &M, st));
    if (M < 1) {
        ctx->err = "kf_fit_series: the series yields no snapshot pairs";
        return KF_EINVAL;
    }
    if (scale) {
        std::vector<double> h(2 * ncol);
        KF_CUDA(ctx, cudaMemcpyAsync(h.data(), d_scale, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        for (int c = 0; c < n; ++c) {
            if (scale->y_offset) scale->y_offset[c] = h[c];
            if (scale->y_factor) scale->y_factor[c] = h[ncol + c];
        }
        for (int c = 0; c < m; ++c) {
            if (scale->u_offset) scale->u_offset[c] = h[n + c];
            if (scale->u_factor) scale->u_factor[c] = h[ncol + n + c];
        }
        scale->M = M;
    }
    const size_t nz = (size_t)M * nzeta, nu = (size_t)M * m;
    KF_CUDA(ctx, ctx->d_in.ensure((2 * nz + nu + 2) * sizeof(double)));
    double* d_alpha = ctx->d_in.as<double>();
    double* d_beta = d_alpha + nz;
    double* d_uo = d_beta + nz;
    KF_TRY(kf_pp_pairs(ctx, d_t, d_y, d_u, d_scale, T, n, m, nd, M, d_alpha, d_beta, d_uo, st));
    kf_problem dp{};
    dp.M = M; dp.nzeta = nzeta; dp.m = m; dp.model = ser->model;
    dp.alpha = d_alpha; dp.beta = d_beta; dp.u = d_uo; dp.pc_cols = ser->pc_cols;
    KF_TRY(check_problem(ctx, &dp));
    return kf_fit_device_pairs(ctx, &dp, solve, out, t0, false, 0, 0);
}

int kf_fit_batch(kf_ctx* ctx, int nprob, const kf_basis* const* bases, const kf_problem* probs, const kf_solve* solves, kf_result* outs) {
    if (!ctx) return KF_EINVAL;
    if (nprob < 0 || (nprob && (!bases || !probs || !solves || !outs))) {
        ctx->err = "kf_fit_batch: bases, probs, solves, outs required";
        return KF_EINVAL;
    }
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<int> small, rest;
    for (int i = 0; i < nprob; ++i) {
        int N = 0, P = 0;
        if (!bases[i]) { ctx->err = "kf_fit_batch: NULL basis"; return KF_EINVAL; }
        KF_TRY(kf_basis_dims(bases[i], probs[i].model, probs[i].m, nullptr, &N, &P));
        // least-squares fits, and QP fits without pinned columns / forced shift: if the LS solution turns out to lie inside
        // every budget (evaluate_rand_models.m fits its nonlinear models with lasso = 4, usually inactive) it is the QP answer too
        const bool solve_ok = solves[i].least_squares ||
                              (solves[i].nt >= 1 && solves[i].t && !solves[i].delay_constraint && solves[i].psd_shift != KF_PSD_ALWAYS);
        const bool ok = solve_ok && P <= 32 && probs[i].pc_cols == 0 && probs[i].nw == 0 && bases[i]->n_pcs == 0 && probs[i].M > 0 && probs[i].alpha &&
                        probs[i].beta && (probs[i].m == 0 || probs[i].u) && !outs[i].G && !outs[i].C && !outs[i].Px && !outs[i].Py &&
                        (probs[i].model == KF_LINEAR || probs[i].model == KF_BILINEAR || probs[i].model == KF_NONLINEAR) &&
                        bases[i]->nv == probs[i].nzeta + (probs[i].model == KF_NONLINEAR ? probs[i].m : 0);
        (ok ? small : rest).push_back(i);
    }
    // concurrent part, in groups that bound the staging memory
    const size_t group = 2048;
    std::vector<int> accepted(group);
    for (size_t g0 = 0; g0 < small.size(); g0 += group) {
        const int n = (int)std::min(group, small.size() - g0);
        KF_TRY(kf_fit_batch_small(ctx, nprob, bases, probs, solves, outs, small.data() + g0, n, accepted.data()));
        for (int w = 0; w < n; ++w)
            if (!accepted[w]) rest.push_back(small[g0 + w]);
    }
    for (int i : rest) KF_TRY(kf_fit(ctx, bases[i], &probs[i], &solves[i], &outs[i]));
    return KF_OK;
}

int kf_rollout(kf_ctx* ctx, const kf_basis* basis, int nmodels, const kf_model* models, int ntrials, const int* T,
               const double* const* zeta0, const double* const* u, int nout, double* const* ysim) {
    if (!ctx) return KF_EINVAL;
    if (!models || nmodels <= 0 || ntrials <= 0 || !T || !zeta0 || !ysim || (models[0].m > 0 && !u)) {
        ctx->err = "kf_rollout: models, T, zeta0, u, ysim required";
        return KF_EINVAL;
    }
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    KF_TRY(prepare_program(ctx, basis));
    ctx->lay.valid = false;
    return kf_rollout_impl(ctx, nmodels, models, ntrials, T, zeta0, u, nout, ysim);
}

int kf_set_qp_partition(kf_ctx* ctx, int col_lo, int col_hi, kf_allreduce_fn allreduce, void* user) {
    if (!ctx) return KF_EINVAL;
    if (col_hi > col_lo && (col_lo < 0 || !allreduce)) {
        ctx->err = "kf_set_qp_partition: col_lo >= 0 and an all-reduce hook are required";
        return KF_EINVAL;
    }
    ctx->qp_lo = col_hi > col_lo ? col_lo : 0;
    ctx->qp_hi = col_hi > col_lo ? col_hi : 0;
    ctx->qp_allreduce = col_hi > col_lo ? allreduce : nullptr;
    ctx->qp_user = user;
    return KF_OK;
}

int kf_mldivide(kf_ctx* ctx, long long M, int P, int Pc, const double* A, const double* B, double* X, int* perm, int* rank) {
    if (!ctx) return KF_EINVAL;
    if (M <= 0 || P <= 0 || Pc <= 0 || !A || !B || !X) {
        ctx->err = "kf_mldivide: M, P, Pc > 0 and A, B, X required";
        return KF_EINVAL;
    }
    KF_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int Pp = (int)kf_roundup(P, KF_BM);
    const long long ldq = kf_qr_ld(M);
    KF_CUDA(ctx, ctx->d_qr.ensure((size_t)ldq * (P + Pc) * sizeof(double)));
    KF_CUDA(ctx, ctx->d_K2.ensure((size_t)Pp * Pc * sizeof(double)));
    KF_CUDA(ctx, ctx->d_misc.ensure(1024 + (size_t)Pp * sizeof(int)));
    double* AB = ctx->d_qr.as<double>();
    if (ldq > M) KF_CUDA(ctx, cudaMemset2DAsync(AB + M, (size_t)ldq * sizeof(double), 0, (size_t)(ldq - M) * sizeof(double), (size_t)(P + Pc), st));
    KF_CUDA(ctx, cudaMemcpy2DAsync(AB, (size_t)ldq * sizeof(double), A, (size_t)M * sizeof(double), (size_t)M * sizeof(double), P, cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpy2DAsync(AB + (size_t)ldq * P, (size_t)ldq * sizeof(double), B, (size_t)M * sizeof(double), (size_t)M * sizeof(double), Pc,
                                   cudaMemcpyHostToDevice, st));
    int* d_perm = reinterpret_cast<int*>(ctx->d_misc.as<char>() + 1024);
    int r = 0;
    double minp = 0, maxp = 0;
    KF_TRY(kf_solve_qr_ls(ctx, M, P, Pc, AB, ldq, ctx->d_K2.as<double>(), Pp, d_perm, &r, &minp, &maxp, st));
    KF_CUDA(ctx, cudaMemcpy2DAsync(X, (size_t)P * sizeof(double), ctx->d_K2.p, (size_t)Pp * sizeof(double), (size_t)P * sizeof(double), Pc,
                                   cudaMemcpyDeviceToHost, st));
    if (perm) KF_CUDA(ctx, cudaMemcpyAsync(perm, d_perm, sizeof(int) * P, cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    if (rank) *rank = r;
    return KF_OK;
}

int kf_counters(kf_ctx* ctx, double* dmma_flops, long long* launches, int reset) {
    if (!ctx) return KF_EINVAL;
    if (dmma_flops) *dmma_flops = ctx->dmma_flops;
    if (launches) *launches = ctx->launches;
    if (reset) {
        ctx->dmma_flops = 0;
        ctx->launches = 0;
    }
    return KF_OK;
}

int kf_engine_info(kf_ctx* ctx, int* engine, double* i8_ops, double* i8_ops_per_launch, long long* gram_launches, int* gram_sampled, int reset) {
    if (!ctx) return KF_EINVAL;
    if (engine) *engine = ctx->last_engine;
    if (i8_ops) *i8_ops = ctx->i8_ops;
    if (i8_ops_per_launch) *i8_ops_per_launch = ctx->i8_ops_per_launch;
    if (gram_launches) *gram_launches = ctx->last_gram_launches;
    if (gram_sampled) *gram_sampled = ctx->last_gram_sampled;
    if (reset) ctx->i8_ops = 0;
    return KF_OK;
}

int kf_last_times(kf_ctx* ctx, double* lift_gram_ms, double* gram_kernel_ms, double* solve_ms) {
    if (!ctx) return KF_EINVAL;
    if (lift_gram_ms) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]) == cudaSuccess) ctx->last_lift_gram_ms = ms;
        *lift_gram_ms = ctx->last_lift_gram_ms + ctx->last_refine_ms;
    }
    if (gram_kernel_ms) *gram_kernel_ms = ctx->last_gram_kernel_ms;
    if (solve_ms) *solve_ms = ctx->last_solve_ms;
    return KF_OK;
}

int kf_set_option(kf_ctx* ctx, const char* name, double value) {
    if (!ctx || !name) return KF_EINVAL;
    const std::string n(name);
    if (n == "chunk") ctx->opt_chunk = (int)value;
    else if (n == "splitk") ctx->opt_splitk = (int)value;
    else if (n == "overlap") ctx->opt_overlap = (int)value;
    else if (n == "panel_mb") ctx->opt_panel_mb = value;
    else if (n == "profile") ctx->opt_profile = (int)value;
    else if (n == "tma") ctx->opt_tma = (int)value;
    else if (n == "qr_max_gb") ctx->opt_qr_max_gb = value;
    else if (n == "qp_method") ctx->opt_qp_method = (int)value;
    else if (n == "as_ws_gb") ctx->opt_as_ws_gb = value;
    else if (n == "as_frac") ctx->opt_as_frac = value;
    else if (n == "as_diag") ctx->opt_as_diag = (int)value;
    else if (n == "as_skip") ctx->opt_as_skip = (int)value;
    else if (n == "as_level") ctx->opt_as_level = (int)value;
    else if (n == "lift_tile") ctx->opt_lift_tile = (int)value;
    else if (n == "lift_ls") ctx->opt_lift_ls = (int)value;
    else if (n == "qr_blocked") ctx->opt_qr_blocked = (int)value;
    else if (n == "qr_nb") ctx->opt_qr_nb = (int)value;
    else if (n == "qr_ksplit") ctx->opt_qr_ksplit = (int)value;
    else if (n == "lift_wide") ctx->opt_lift_wide = (int)value;
    else if (n == "lift_smem_kb") ctx->opt_lift_smem_kb = value;
    else if (n == "lift_minb") ctx->opt_lift_minb = (int)value;
    else if (n == "lift_panel_fit") ctx->opt_lift_panel_fit = (int)value;
    else if (n == "graphs") ctx->opt_graphs = (int)value;
    else if (n == "gram_engine") ctx->opt_gram_engine = (int)value;
    else if (n == "oz_sym") ctx->opt_oz_sym = (int)value;
    else if (n == "oz_pipes") ctx->opt_oz_pipes = (int)value;
    else if (n == "qp_split") ctx->opt_qp_split = (int)value;
    else if (n == "refine") ctx->opt_refine = (int)value;
    else if (n == "refine_kappa") ctx->opt_refine_kappa = value;
    else if (n == "refine_level_tol") ctx->opt_refine_level_tol = value;
    else if (n == "refine_max") ctx->opt_refine_max = (int)value;
    else {
        ctx->err = "kf_set_option: unknown option " + n;
        return KF_EINVAL;
    }
    return KF_OK;
}

}  // extern "C"
