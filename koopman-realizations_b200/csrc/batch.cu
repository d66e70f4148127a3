// Batched small fits: many independent least-squares Koopman fits with P <= 32 solved concurrently,
// ONE CTA PER PROBLEM, everything on chip (BASELINE config 4: evaluate_rand_models.m:45-144 runs 23 tiny fits
// per random system, hundreds of systems).
//
// Per problem: stream the snapshots in tiles of 256 rows; lift each row (same feature program / op semantics as
// lift.cu) into the shared tile T = [Px | Py] (256 x 2P); fold the tile into the running triangular factor
// R = [R11 | Q'Py] (P x 2P) with structured Householder reflectors (TSQR: the stacked matrix [R; T] is reduced
// column by column, the reflector of column j touches row j of R and the 256 tile rows only).  After the last
// tile, R11'R11 = Px'Px exactly as for a QR of the whole Px, so a final Householder QR WITH column pivoting of the
// small R11 (applied to Q'Py) reproduces mldivide's pivoted basic solution (Ksysid.m:1069) with QR-level accuracy
// (no squaring of the condition number, unlike a batched Gram + Cholesky).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "kf_internal.h"
#include "lift_eval.h"

namespace {

constexpr int BT_ROWS = 256;       // snapshots per tile = threads per CTA
constexpr int BT_PMAX = 32;        // regressor width limit of the batched path
constexpr int BT_W = 2 * BT_PMAX;  // columns of [Px | Py]
constexpr int BT_LD = BT_W + 1;    // padded row of the shared tile

struct KfBatchDesc {
    long long M;
    long long in_off;    // packed inputs: alpha (M x nzeta) | beta (M x nzeta) | u (M x m), doubles
    long long out_off;   // packed outputs: K (P x P), doubles
    int nzeta, m, model;
    int nv, n_full, N, P;
    int op_off, cen_off;
    int pad;
};

struct KfBatchOut {
    int rank;
    int perm[BT_PMAX];
    double r11, minpiv;
    // of the returned K (for the QP branch with an inactive budget, where K is also the QP minimiser):
    double proj2;      // ||Q1'Py||_F^2 = tr(C'K): objective 0.5 tr(K'GK) - tr(C'K) = -0.5 proj2
    double l1;         // ||vec K||_1
    double ginner;     // <G K - C, K>   (rounding level)
    double gmax;       // ||G K - C||_inf
};

__device__ __forceinline__ double warp_sum(double v) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return __shfl_sync(0xffffffffu, v, 0);
}

__global__ void __launch_bounds__(BT_ROWS, 1)
kf_batch_ls_kernel(const KfBatchDesc* __restrict__ descs, const KfOp* __restrict__ ops_all, const double* __restrict__ cen_all,
                   const double* __restrict__ in_all, double* __restrict__ K_all, KfBatchOut* __restrict__ outs) {
    extern __shared__ __align__(16) double sm[];
    double* T = sm;                              // BT_ROWS x BT_LD
    double* R = T + BT_ROWS * BT_LD;             // BT_PMAX x BT_LD
    double* v = R + BT_PMAX * BT_LD;             // BT_ROWS reflector entries
    __shared__ double s_red[8];
    __shared__ double s_beta, s_tau;
    __shared__ int s_skip;

    const KfBatchDesc d = descs[blockIdx.x];
    const KfOp* ops = ops_all + d.op_off;
    const double* cen = cen_all + d.cen_off;
    const double* alpha = in_all + d.in_off;
    const double* beta = alpha + d.M * d.nzeta;
    const double* u = beta + d.M * d.nzeta;
    const int P = d.P, W = 2 * d.P, N = d.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int e = tid; e < BT_PMAX * BT_LD; e += BT_ROWS) R[e] = 0.0;
    __syncthreads();

    for (long long s0 = 0; s0 < d.M; s0 += BT_ROWS) {
        // ---- lift one snapshot per thread into its tile row
        {
            const long long s = s0 + tid;
            double* row = T + tid * BT_LD;
            if (s < d.M) {
                double f[BT_PMAX + 1];
                for (int side = 0; side < 2; ++side) {
                    const double* src = side ? beta : alpha;
                    for (int i = 0; i < d.nv; ++i)
                        f[i] = i < d.nzeta ? src[(long long)i * d.M + s] : u[(long long)(i - d.nzeta) * d.M + s];
                    for (int j = d.nv; j < d.n_full; ++j) f[j] = kf_eval_op(ops[j], d.nv, cen, [&](int k) { return f[k]; });
                    double* dst = row + side * P;
                    for (int j = 0; j < N; ++j) dst[j] = f[j];
                    if (d.model == KF_LINEAR) {            // [psi, u]  (Ksysid.m:1062-1063)
                        for (int i = 0; i < d.m; ++i) dst[N + i] = u[(long long)i * d.M + s];
                    } else if (d.model == KF_BILINEAR) {   // [psi; u_k psi]  (Ksysid.m:510-511)
                        for (int k = 0; k < d.m; ++k) {
                            const double uk = u[(long long)k * d.M + s];
                            for (int j = 0; j < N; ++j) dst[(k + 1) * N + j] = KF_MUL(uk, f[j]);
                        }
                    }
                }
            } else {
                for (int c = 0; c < W; ++c) row[c] = 0.0;
            }
        }
        __syncthreads();
        // ---- fold the tile into R: structured Householder on [R; T], column by column
        for (int j = 0; j < P; ++j) {
            const double x = T[tid * BT_LD + j];
            double sig = warp_sum(x * x);
            if (lane == 0) s_red[warp] = sig;
            __syncthreads();
            if (tid == 0) {
                double sigma = 0.0;
                for (int w = 0; w < 8; ++w) sigma += s_red[w];
                const double a = R[j * BT_LD + j];
                if (sigma == 0.0) {
                    s_skip = 1;
                } else {
                    const double bt = -copysign(sqrt(a * a + sigma), a);
                    s_beta = bt;
                    s_tau = (bt - a) / bt;
                    s_red[0] = 1.0 / (a - bt);
                    s_skip = 0;
                }
            }
            __syncthreads();
            if (s_skip) continue;
            const double scale = s_red[0], tau = s_tau;
            v[tid] = x * scale;
            __syncthreads();
            // warps own columns c = j+1+warp, j+1+warp+8, ...: w = R(j,c) + v'T(:,c); R(j,c) -= tau w; T(:,c) -= tau w v
            for (int c = j + 1 + warp; c < W; c += 8) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < BT_ROWS / 32; ++q) acc = fma(v[lane + 32 * q], T[(lane + 32 * q) * BT_LD + c], acc);
                const double w = tau * (warp_sum(acc) + R[j * BT_LD + c]);
#pragma unroll
                for (int q = 0; q < BT_ROWS / 32; ++q) {
                    const int i = lane + 32 * q;
                    T[i * BT_LD + c] = fma(-w, v[i], T[i * BT_LD + c]);
                }
                if (lane == 0) R[j * BT_LD + c] -= w;
            }
            if (tid == 0) R[j * BT_LD + j] = s_beta;
            __syncthreads();
        }
    }

    // ---- final step, one warp: Householder QR with column pivoting of R11 (P x P), reflectors applied to Q'Py;
    //      lane = row of R.  Rank by MATLAB's max(size(A)) * eps(|R11|); basic solution by back-substitution.
    if (warp == 0) {
        KfBatchOut* out = outs + blockIdx.x;
        double* Kout = K_all + d.out_off;
        __shared__ int perm[BT_PMAX];
        __shared__ double vn[BT_PMAX];
        if (lane < P) perm[lane] = lane;
        // squared column norms of R11
        for (int c = 0; c < P; ++c) {
            const double xv = lane < P ? R[lane * BT_LD + c] : 0.0;
            const double nn = warp_sum(xv * xv);
            if (lane == 0) vn[c] = nn;
        }
        __syncwarp();
        int rank = 0;
        double r11 = 0.0, minpiv = 0.0, tol = 0.0;
        for (int j = 0; j < P; ++j) {
            // pivot: largest remaining norm (first index on ties)
            double best = -1.0;
            int bi = P;
            for (int c = j + lane; c < P; c += 32)
                if (vn[c] > best) { best = vn[c]; bi = c; }
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, off);
                const int oi = __shfl_down_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            best = __shfl_sync(0xffffffffu, best, 0);
            bi = __shfl_sync(0xffffffffu, bi, 0);
            const double nrm = sqrt(best > 0.0 ? best : 0.0);
            if (j == 0) {
                r11 = nrm;
                const double e = nrm > 0.0 ? (__longlong_as_double(__double_as_longlong(nrm) + 1) - nrm) : 0.0;
                tol = (double)(d.M > P ? d.M : P) * e;
            }
            if (!(nrm > tol)) break;
            rank = j + 1;
            minpiv = nrm;
            if (bi != j) {   // swap columns j <-> bi of R11
                if (lane < P) {
                    const double a = R[lane * BT_LD + j];
                    R[lane * BT_LD + j] = R[lane * BT_LD + bi];
                    R[lane * BT_LD + bi] = a;
                }
                if (lane == 0) {
                    const int a = perm[j]; perm[j] = perm[bi]; perm[bi] = a;
                    vn[bi] = vn[j];
                }
                __syncwarp();
            }
            // reflector from rows j..P-1 of column j
            const double xj = (lane > j && lane < P) ? R[lane * BT_LD + j] : 0.0;
            const double sigma = warp_sum(xj * xj);
            const double a = R[j * BT_LD + j];
            double tau = 0.0, bt = a, scale = 0.0;
            if (sigma > 0.0) {
                bt = -copysign(sqrt(a * a + sigma), a);
                tau = (bt - a) / bt;
                scale = 1.0 / (a - bt);
            }
            const double vj = (lane == j) ? 1.0 : xj * scale;   // zero for lanes < j or >= P
            __syncwarp();
            for (int c = j + 1; c < W; ++c) {
                const double rc = lane < P ? R[lane * BT_LD + c] : 0.0;
                const double w = tau * warp_sum(vj * rc);
                if (lane >= j && lane < P) R[lane * BT_LD + c] = rc - w * vj;
            }
            if (lane == j) R[j * BT_LD + j] = bt;
            if (lane > j && lane < P) R[lane * BT_LD + j] = 0.0;
            __syncwarp();
            // remaining norms of the trailing columns (rows > j), recomputed exactly
            for (int c = j + 1; c < P; ++c) {
                const double xv = (lane > j && lane < P) ? R[lane * BT_LD + c] : 0.0;
                const double nn = warp_sum(xv * xv);
                if (lane == 0) vn[c] = nn;
            }
            __syncwarp();
        }
        // basic solution: X(perm[0:rank], c) = R11(0:rank,0:rank)^-1 (Q'Py)(0:rank, c); lane = right-hand side c
        for (int e = lane; e < P * P; e += 32) Kout[e] = 0.0;
        __syncwarp();
        for (int c = lane; c < P; c += 32) {
            double xs[BT_PMAX];
            for (int i = rank - 1; i >= 0; --i) {
                double acc = R[i * BT_LD + P + c];
                for (int k = i + 1; k < rank; ++k) acc = fma(-R[i * BT_LD + k], xs[k], acc);
                xs[i] = acc / R[i * BT_LD + i];
            }
            for (int i = 0; i < rank; ++i) Kout[(long long)c * P + perm[i]] = xs[i];
        }
        // objective pieces and the gradient G K - C = R11'(R11 K - Q1'Py) of column `lane` (pivoted order)
        double proj2 = 0.0, l1 = 0.0, ginner = 0.0, gmax = 0.0;
        if (lane < P) {
            const int c = lane;
            double xs[BT_PMAX], es[BT_PMAX];
            for (int i = rank - 1; i >= 0; --i) {       // the same recurrence as above (bit-identical xs)
                double acc = R[i * BT_LD + P + c];
                for (int k = i + 1; k < rank; ++k) acc = fma(-R[i * BT_LD + k], xs[k], acc);
                xs[i] = acc / R[i * BT_LD + i];
            }
            for (int i = 0; i < rank; ++i) {
                const double rhs = R[i * BT_LD + P + c];
                double acc = -rhs;
                for (int k = i; k < rank; ++k) acc = fma(R[i * BT_LD + k], xs[k], acc);
                es[i] = acc;
                proj2 = fma(rhs, rhs, proj2);
                l1 += fabs(xs[i]);
            }
            for (int k = 0; k < rank; ++k) {
                double g = 0.0;
                for (int i = 0; i <= k; ++i) g = fma(R[i * BT_LD + k], es[i], g);
                ginner = fma(g, xs[k], ginner);
                gmax = fmax(gmax, fabs(g));
            }
        }
        proj2 = warp_sum(proj2);
        l1 = warp_sum(l1);
        ginner = warp_sum(ginner);
        for (int off = 16; off > 0; off >>= 1) gmax = fmax(gmax, __shfl_down_sync(0xffffffffu, gmax, off));
        if (lane == 0) {
            out->rank = rank;
            out->r11 = r11;
            out->minpiv = minpiv;
            out->proj2 = proj2;
            out->l1 = l1;
            out->ginner = ginner;
            out->gmax = gmax;
        }
        if (lane < BT_PMAX) out->perm[lane] = lane < P ? perm[lane] : -1;
    }
}

}  // namespace

// nprob independent LS fits; problems with P > 32 (or non-LS solves) are not handled here (caller falls back to kf_fit)
// (solves[ip].least_squares == 0: the QP branch — accepted[w] = 1 only if the LS solution has full rank and lies inside every
// budget, in which case it is the QP minimiser too (Ksysid.m:1135-1137 inactive, no 1e-6 I shift); otherwise the caller
// sends the problem down the general path)
int kf_fit_batch_small(kf_ctx* ctx, int nprob, const kf_basis* const* bases, const kf_problem* probs, const kf_solve* solves, kf_result* outs,
                       const int* which, int nwhich, int* accepted) {
    cudaStream_t st = ctx->stream;
    std::vector<KfBatchDesc> descs;
    std::vector<KfOp> ops;
    std::vector<double> cen;
    std::vector<double> inputs;
    std::map<std::pair<const void*, long long>, long long> seen;   // shared snapshot sets are uploaded once
    long long out_off = 0;
    std::string err;
    (void)nprob;
    for (int w = 0; w < nwhich; ++w) {
        const int ip = which[w];
        KfProgram prog;
        int rc = kf_build_program(bases[ip], prog, err);
        if (rc) { ctx->err = err; return rc; }
        const kf_problem& pr = probs[ip];
        KfBatchDesc d{};
        d.M = pr.M; d.nzeta = pr.nzeta; d.m = pr.m; d.model = pr.model;
        d.nv = prog.nv; d.n_full = prog.n_full(); d.N = prog.N();
        d.P = kf_regressor_width(pr.model, d.N, pr.m);
        d.op_off = (int)ops.size();
        d.cen_off = (int)cen.size();
        ops.insert(ops.end(), prog.ops.begin(), prog.ops.end());
        cen.insert(cen.end(), prog.centres.begin(), prog.centres.end());
        auto key = std::make_pair((const void*)pr.alpha, pr.M);
        auto it = seen.find(key);
        if (it == seen.end()) {
            d.in_off = (long long)inputs.size();
            seen[key] = d.in_off;
            inputs.insert(inputs.end(), pr.alpha, pr.alpha + pr.M * pr.nzeta);
            inputs.insert(inputs.end(), pr.beta, pr.beta + pr.M * pr.nzeta);
            if (pr.m) inputs.insert(inputs.end(), pr.u, pr.u + pr.M * pr.m);
        } else {
            d.in_off = it->second;
        }
        d.out_off = out_off;
        out_off += (long long)d.P * d.P;
        descs.push_back(d);
    }
    if (cen.empty()) cen.push_back(0.0);
    KF_CUDA(ctx, ctx->d_in.ensure(inputs.size() * sizeof(double)));
    KF_CUDA(ctx, ctx->d_ops.ensure(ops.size() * sizeof(KfOp)));
    KF_CUDA(ctx, ctx->d_centres.ensure(cen.size() * sizeof(double)));
    KF_CUDA(ctx, ctx->d_tasks[0].ensure(descs.size() * sizeof(KfBatchDesc)));
    KF_CUDA(ctx, ctx->d_K.ensure((size_t)std::max<long long>(out_off, 1) * sizeof(double)));
    KF_CUDA(ctx, ctx->d_tmp.ensure(descs.size() * sizeof(KfBatchOut)));
    KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_in.p, inputs.data(), inputs.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_ops.p, ops.data(), ops.size() * sizeof(KfOp), cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_centres.p, cen.data(), cen.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(ctx->d_tasks[0].p, descs.data(), descs.size() * sizeof(KfBatchDesc), cudaMemcpyHostToDevice, st));
    const size_t smem = (size_t)(BT_ROWS * BT_LD + BT_PMAX * BT_LD + BT_ROWS) * sizeof(double);
    KF_CUDA(ctx, kf_ensure_smem(ctx, kf_batch_ls_kernel, smem));
    KF_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
    kf_batch_ls_kernel<<<(unsigned)descs.size(), BT_ROWS, smem, st>>>(ctx->d_tasks[0].as<KfBatchDesc>(), ctx->d_ops.as<KfOp>(),
                                                                      ctx->d_centres.as<double>(), ctx->d_in.as<double>(),
                                                                      ctx->d_K.as<double>(), ctx->d_tmp.as<KfBatchOut>());
    KF_CUDA(ctx, cudaGetLastError());
    KF_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
    ctx->launches += 1;
    std::vector<double> Kh((size_t)std::max<long long>(out_off, 1));
    std::vector<KfBatchOut> oh(descs.size());
    KF_CUDA(ctx, cudaMemcpyAsync(Kh.data(), ctx->d_K.p, (size_t)out_off * sizeof(double), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaMemcpyAsync(oh.data(), ctx->d_tmp.p, descs.size() * sizeof(KfBatchOut), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    KF_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    ctx->last_solve_ms = ms;
    for (int w = 0; w < nwhich; ++w) {
        const int ip = which[w];
        const KfBatchDesc& d = descs[w];
        kf_result& o = outs[ip];
        const kf_solve& sv = solves[ip];
        accepted[w] = 1;
        if (!sv.least_squares) {
            double tmin = sv.nt > 0 ? sv.t[0] : 0.0;
            for (int q = 1; q < sv.nt; ++q) tmin = std::min(tmin, sv.t[q]);
            if (oh[w].rank < d.P || !(oh[w].l1 <= tmin)) { accepted[w] = 0; continue; }
        }
        std::memset(&o.info, 0, sizeof(o.info));
        const int nt = sv.least_squares ? 1 : std::max(sv.nt, 1);
        for (int q = 0; q < nt; ++q) {
            if (o.K) std::memcpy(o.K + (size_t)q * d.P * d.P, Kh.data() + d.out_off, (size_t)d.P * d.P * sizeof(double));
            if (!sv.least_squares) {
                if (o.objective) o.objective[q] = -0.5 * oh[w].proj2;
                if (o.l1norm) o.l1norm[q] = oh[w].l1;
                if (o.qp_iters) o.qp_iters[q] = 0;
                if (o.qp_gap) o.qp_gap[q] = std::max(0.0, oh[w].ginner + sv.t[q] * oh[w].gmax);
            }
        }
        if (o.perm) std::memcpy(o.perm, oh[w].perm, sizeof(int) * d.P);
        o.info.rank = oh[w].rank;
        o.info.ls_method_used = KF_LS_QR;
        o.info.max_pivot = oh[w].r11;
        o.info.min_pivot = oh[w].minpiv;
        o.info.passes = 1;
        o.info.t_solve_ms = ms;
    }
    ctx->lay.valid = false;   // d_ops / d_tasks were reused
    return KF_OK;
}
