// Open-loop validation rollouts on the GPU (SURVEY §8f "next #2"): val_model / val_BLmodel / val_NLmodel
// (Ksysid.m:1623-1879, discrete time, unloaded).  One CTA per validation trial; the lifted state lives in shared
// memory and every step is a matrix-vector product with A (N x N, column-major: thread i owns row i, so the loads of
// a column are coalesced) plus the input term; the nonlinear model re-lifts [zeta; u] each step with the same
// feature-program evaluator as lift.cu, level by level.
#include <algorithm>
#include <vector>

#include "kf_internal.h"
#include "lift_eval.h"

namespace {

constexpr int RO_MAX_THREADS = 1024;
#define RO_THREADS ((int)blockDim.x)

struct RolloutArgs {
    const KfOp* ops; const double* centres; const double* pcs; const int* order; const int* level_start;
    int nlevels, nv, n_full, n_pcs, N;
    int model, n, m, nzeta, nout, ntrials;
    const double* A; const double* B; const double* F;      // candidate c at A + c * strideA etc.
    long long strideA, strideB, strideF, y_total;
    const int* T; const long long* u_off; const long long* y_off; const double* zeta0;   // per trial
    const double* u_all; double* y_all;
};

// lift one point v (nv values in `v`) into z (N values); f is scratch of n_full doubles; all in shared memory
__device__ void lift_point(const RolloutArgs& a, const double* v, double* f, double* z) {
    const int tid = threadIdx.x;
    for (int l = 0; l < a.nlevels; ++l) {
        const int first = a.level_start[l], count = a.level_start[l + 1] - first;
        for (int e = tid; e < count; e += RO_THREADS) {
            const int j = a.order[first + e];
            const KfOp op = a.ops[j];
            auto feat = [&](int k) -> double { return k < a.nv ? v[k] : f[k]; };
            f[j] = (op.kind == KF_OP_VAR) ? v[op.a] : kf_eval_op(op, a.nv, a.centres, feat);
        }
        __syncthreads();
    }
    if (a.n_pcs == 0) {
        for (int j = tid; j < a.n_full; j += RO_THREADS) z[j] = f[j];
    } else {   // econ lift [v; pcs' f; 1]  (Ksysid.m:1614-1618)
        for (int i = tid; i < a.nv; i += RO_THREADS) z[i] = v[i];
        for (int c = tid; c < a.n_pcs; c += RO_THREADS) {
            const double* pc = a.pcs + (size_t)c * a.n_full;
            double acc = 0.0;
            for (int j = 0; j < a.n_full; ++j) acc = fma(pc[j], f[j], acc);
            z[a.nv + c] = acc;
        }
        if (tid == 0) z[a.nv + a.n_pcs] = 1.0;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RO_MAX_THREADS) kf_rollout_kernel(const RolloutArgs a) {
    extern __shared__ __align__(16) double ro_smem[];
    const int trial = blockIdx.x, cand = blockIdx.y, tid = threadIdx.x;
    const int N = a.N, T = a.T[trial], nout = a.nout;
    const double* __restrict__ A = a.A + cand * a.strideA;
    const double* __restrict__ B = a.B + cand * a.strideB;
    const double* __restrict__ F = a.F + cand * a.strideF;
    double* z = ro_smem;                 // lifted state (linear, bilinear) or lifted [zeta; u] (nonlinear)
    double* zn = z + N;
    double* f = zn + N;                  // n_full scratch
    double* v = f + a.n_full;            // nv point being lifted
    const double* u = a.u_all + a.u_off[trial];        // T x m column-major
    double* y = a.y_all + cand * a.y_total + a.y_off[trial];      // T x nout column-major
    const double* zeta0 = a.zeta0 + (size_t)trial * a.nzeta;

    if (a.model != KF_NONLINEAR) {
        for (int i = tid; i < a.nzeta; i += RO_THREADS) v[i] = zeta0[i];
        __syncthreads();
        lift_point(a, v, f, z);                         // z0 = lift.econ_full(zeta0)   (Ksysid.m:1674, 1768)
        for (int i = tid; i < nout; i += RO_THREADS) y[i * (long long)T] = z[i];      // y = C z, C = [I_n 0]; zeta, z: more rows
        for (int j = 0; j < T - 1; ++j) {
            for (int i = tid; i < N; i += RO_THREADS) {
                double acc = 0.0;
                for (int k = 0; k < N; ++k) acc = fma(A[(size_t)k * N + i], z[k], acc);
                if (a.model == KF_LINEAR) {             // z+ = A z + B u   (Ksysid.m:1685)
                    for (int q = 0; q < a.m; ++q) acc = fma(B[(size_t)q * N + i], u[(size_t)q * T + j], acc);
                } else {                                // z+ = A z + sum_q u_q B_q z   (Beta(z) u, Ksysid.m:1783, 1288-1289)
                    for (int q = 0; q < a.m; ++q) {
                        const double uq = u[(size_t)q * T + j];
                        const double* Bq = B + (size_t)q * N * N;
                        double bacc = 0.0;
                        for (int k = 0; k < N; ++k) bacc = fma(Bq[(size_t)k * N + i], z[k], bacc);
                        acc = fma(uq, bacc, acc);
                    }
                }
                zn[i] = acc;
            }
            __syncthreads();
            for (int i = tid; i < N; i += RO_THREADS) z[i] = zn[i];
            for (int i = tid; i < nout; i += RO_THREADS) y[i * (long long)T + j + 1] = zn[i];
            __syncthreads();
        }
    } else {
        double* zeta = zn;                               // current zeta (nzeta values) kept in the zn area
        for (int i = tid; i < a.nzeta; i += RO_THREADS) zeta[i] = zeta0[i];
        __syncthreads();
        for (int i = tid; i < nout; i += RO_THREADS) y[i * (long long)T] = zeta[i];
        for (int j = 0; j < T - 1; ++j) {
            for (int i = tid; i < a.nv; i += RO_THREADS) v[i] = i < a.nzeta ? zeta[i] : u[(size_t)(i - a.nzeta) * T + j];
            __syncthreads();
            lift_point(a, v, f, z);                      // psi([zeta; u])
            const int warp = tid >> 5, lane = tid & 31;
            for (int i = warp; i < a.nzeta; i += RO_THREADS / 32) {     // zeta+ = F psi   (Ksysid.m:1860, 1329)
                double acc = 0.0;
                for (int k = lane; k < N; k += 32) acc = fma(F[(size_t)k * a.nzeta + i], z[k], acc);
                for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
                if (lane == 0) zeta[i] = acc;
            }
            __syncthreads();
            for (int i = tid; i < nout; i += RO_THREADS) y[i * (long long)T + j + 1] = zeta[i];
            __syncthreads();
        }
    }
}

}  // namespace

int kf_rollout_impl(kf_ctx* ctx, int nmodels, const kf_model* mdls, int ntrials, const int* T, const double* const* zeta0,
                    const double* const* u, int nout, double* const* ysim) {
    cudaStream_t st = ctx->stream;
    const KfProgram& p = ctx->prog;
    const int N = p.N();
    const kf_model* mdl = &mdls[0];
    const bool nl = mdl->model == KF_NONLINEAR;
    for (int c = 0; c < nmodels; ++c) {
        const kf_model& q = mdls[c];
        if (q.model != mdl->model || q.n != mdl->n || q.m != mdl->m || q.nzeta != mdl->nzeta || q.N != mdl->N) {
            ctx->err = "kf_rollout: all candidate models must share model type and dimensions";
            return KF_EINVAL;
        }
        if ((nl && !q.F) || (!nl && (!q.A || (q.m > 0 && !q.B)))) {
            ctx->err = "kf_rollout: A, B (linear, bilinear) or F (nonlinear) required";
            return KF_EINVAL;
        }
    }
    if (mdl->N != N) {
        ctx->err = "kf_rollout: kf_model.N does not match the dictionary";
        return KF_EINVAL;
    }
    const int nv = mdl->nzeta + (nl ? mdl->m : 0);
    if (p.nv != nv) {
        ctx->err = "kf_rollout: kf_basis.nv must equal nzeta (linear, bilinear) or nzeta+m (nonlinear)";
        return KF_EINVAL;
    }
    if (nout <= 0) nout = mdl->n;
    if (nout > (nl ? mdl->nzeta : N)) {
        ctx->err = "kf_rollout: nout exceeds the state dimension (N, or nzeta for the nonlinear model)";
        return KF_EINVAL;
    }
    // pack per-trial inputs
    std::vector<long long> u_off(ntrials), y_off(ntrials);
    long long nu = 0, ny = 0;
    for (int k = 0; k < ntrials; ++k) {
        if (T[k] < 1) { ctx->err = "kf_rollout: every trial needs T >= 1"; return KF_EINVAL; }
        u_off[k] = nu; y_off[k] = ny;
        nu += (long long)T[k] * mdl->m;
        ny += (long long)T[k] * nout;
    }
    std::vector<double> hu((size_t)std::max<long long>(nu, 1)), hz((size_t)ntrials * mdl->nzeta);
    for (int k = 0; k < ntrials; ++k) {
        if (mdl->m) std::copy(u[k], u[k] + (size_t)T[k] * mdl->m, hu.begin() + u_off[k]);
        std::copy(zeta0[k], zeta0[k] + mdl->nzeta, hz.begin() + (size_t)k * mdl->nzeta);
    }
    const size_t nA = nl ? 0 : (size_t)N * N;
    const size_t nB = mdl->model == KF_LINEAR ? (size_t)N * mdl->m : (mdl->model == KF_BILINEAR ? (size_t)N * N * mdl->m : 0);
    const size_t nF = nl ? (size_t)mdl->nzeta * N : 0;
    // device layout in d_qr: A[c] | B[c] | F[c] | u | zeta0 | y[c] ; d_tmp: u_off | y_off | T | level_start
    const size_t nd = (nA + nB + nF) * nmodels + hu.size() + hz.size() + (size_t)ny * nmodels;
    KF_CUDA(ctx, ctx->d_qr.ensure(nd * sizeof(double)));
    double* dA = ctx->d_qr.as<double>();
    double* dB = dA + nA * nmodels;
    double* dF = dB + nB * nmodels;
    double* du = dF + nF * nmodels;
    double* dz = du + hu.size();
    double* dy = dz + hz.size();
    for (int c = 0; c < nmodels; ++c) {
        if (nA) KF_CUDA(ctx, cudaMemcpyAsync(dA + nA * c, mdls[c].A, nA * sizeof(double), cudaMemcpyHostToDevice, st));
        if (nB) KF_CUDA(ctx, cudaMemcpyAsync(dB + nB * c, mdls[c].B, nB * sizeof(double), cudaMemcpyHostToDevice, st));
        if (nF) KF_CUDA(ctx, cudaMemcpyAsync(dF + nF * c, mdls[c].F, nF * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    KF_CUDA(ctx, cudaMemcpyAsync(du, hu.data(), hu.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(dz, hz.data(), hz.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    const int nlev = (int)ctx->level_start.size() - 1;
    const size_t ni = (size_t)ntrials + ctx->level_start.size();
    KF_CUDA(ctx, ctx->d_tmp.ensure(ni * sizeof(int) + 2 * (size_t)ntrials * sizeof(long long) + 64));
    long long* d_uoff = ctx->d_tmp.as<long long>();
    long long* d_yoff = d_uoff + ntrials;
    int* d_T = reinterpret_cast<int*>(d_yoff + ntrials);
    int* d_lev = d_T + ntrials;
    KF_CUDA(ctx, cudaMemcpyAsync(d_uoff, u_off.data(), sizeof(long long) * ntrials, cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(d_yoff, y_off.data(), sizeof(long long) * ntrials, cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(d_T, T, sizeof(int) * ntrials, cudaMemcpyHostToDevice, st));
    KF_CUDA(ctx, cudaMemcpyAsync(d_lev, ctx->level_start.data(), sizeof(int) * ctx->level_start.size(), cudaMemcpyHostToDevice, st));

    RolloutArgs a{};
    a.ops = ctx->d_ops.as<KfOp>(); a.centres = ctx->d_centres.as<double>(); a.pcs = ctx->d_pcs.as<double>();
    a.order = ctx->d_order.as<int>(); a.level_start = d_lev; a.nlevels = nlev;
    a.nv = p.nv; a.n_full = p.n_full(); a.n_pcs = p.n_pcs; a.N = N;
    a.model = mdl->model; a.n = mdl->n; a.m = mdl->m; a.nzeta = mdl->nzeta; a.nout = nout; a.ntrials = ntrials;
    a.A = dA; a.B = dB; a.F = dF;
    a.strideA = (long long)nA; a.strideB = (long long)nB; a.strideF = (long long)nF; a.y_total = ny;
    a.T = d_T; a.u_off = d_uoff; a.y_off = d_yoff; a.zeta0 = dz; a.u_all = du; a.y_all = dy;
    const size_t smem = (size_t)(2 * N + p.n_full() + p.nv + 8) * sizeof(double);
    if (smem > 227 * 1024) {
        ctx->err = "kf_rollout: lifted state does not fit in shared memory (N + n_full too large)";
        return KF_EINVAL;
    }
    KF_CUDA(ctx, kf_ensure_smem(ctx, kf_rollout_kernel, smem));
    const int width = std::max(N, p.n_full());
    const int threads = std::min(RO_MAX_THREADS, std::max(128, (width + 31) / 32 * 32));
    kf_rollout_kernel<<<dim3(ntrials, nmodels), threads, smem, st>>>(a);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    std::vector<double> hy((size_t)std::max<long long>(ny * nmodels, 1));
    KF_CUDA(ctx, cudaMemcpyAsync(hy.data(), dy, (size_t)ny * nmodels * sizeof(double), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    for (int c = 0; c < nmodels; ++c)
        for (int k = 0; k < ntrials; ++k) {
            const double* src = hy.data() + (size_t)c * ny + y_off[k];
            std::copy(src, src + (size_t)T[k] * nout, ysim[(size_t)c * ntrials + k]);
        }
    return KF_OK;
}
