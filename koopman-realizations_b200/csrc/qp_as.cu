// Exact active-set solver for the L1-ball QP of Ksysid.solve_KoopmanQP (Ksysid.m:1095-1176)
//
//     min_K  0.5 tr(K' G K) - tr(C' K)    s.t.  ||vec K||_1 <= t
//
// for Gram matrices on which first-order methods stall (config 3a: fourier degree 4, P = 1464, cond(G) ~ 2e13 —
// Jacobi-PCG reaches only 3e-4 relative objective error after 1500 iterations, coordinate descent needs ~1e4 sweeps).
//
// KKT: one multiplier lam >= 0 for the budget; on the support S_j of column j with signs s_j,
//     G_SS k_S = c_S - lam s_S,   sum_j s_j' k_j = t,   |(G k_j - c_j)_i| <= lam off the support.
// Primal-dual active-set iteration (every step EXACT on the current sign pattern):
//   1. per column j, ONE CTA: left-looking blocked Cholesky of G_SjSj (gathered on the fly from the shared G) with the
//      two right-hand sides [c_S, s_S] carried as extra rows, back-substitution:  a_j = G_SS^-1 c_S,  b_j = G_SS^-1 s_S;
//   2. the multiplier from the budget:  lam = (sum_j s_j'a_j - t) / (sum_j s_j'b_j);   K = A - lam B;
//   3. grad = G K - C with the DMMA GEMM; entries whose sign flipped leave the support, entries off the support with
//      |grad| > lam enter it with sign -sign(grad).
// It stops when nothing enters or leaves: K is then the exact minimiser (Frank-Wolfe gap at rounding level).  A lasso
// vector is traversed in ascending order of t, each budget warm-started from the support of the previous one.
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <numeric>

#include "kf_internal.h"

namespace {

constexpr int AS_THREADS = 256;
constexpr int AS_NB = 32;      // block-column width of the factorisation
constexpr int AS_KC = 16;      // contraction elements per pipeline stage
constexpr int AS_NST = 4;      // cp.async stages

// columns a rank works on: [lo, hi) minus the pinned delay columns [skip0, skip1)
struct AsCols {
    int lo, hi, skip0, skip1;
    __host__ __device__ bool on(int c) const { return c >= lo && c < hi && !(c >= skip0 && c < skip1); }
};
constexpr int AS_NL = 16;      // trial multipliers evaluated per pass of the step control

struct AsArgs {
    const double* G; long long ldg;
    const double* C; const double* SG; double* Aout; double* Bout; long long ld;   // P x P, column-major
    const int* idx;            // [column][P]: ascending support indices
    const int* cnt;            // [column]: support size
    const int* cols;           // launch slot -> column
    const long long* ws_off;   // launch slot -> offset (doubles) of its factor in ws
    double* ws;
    double* col_num; double* col_den;   // [column]: s'a, s'b
    int* col_dead;             // [column]: pivots skipped (numerically singular G_SS)
    int P;
    // Level-synchronous mode (few columns per rank: a column gets several CTAs).  mode 0: the whole column in one CTA.  mode 1: ONE
    // block column J0, row tile blockIdx.y + tile_off of it (tile 0 holds the diagonal block; the other tiles are launched after it
    // and read the diagonal factor and the skipped-pivot flags back from global memory).  mode 2: back-substitution + scatter only.
    int mode, J0, tile_off;
    int* dead;                 // mode 1 / 2: [launch slot][P] skipped-pivot flags
};

__device__ __forceinline__ int as_ldl(int n) { return (n + 2 + 1) & ~1; }

// 16-byte asynchronous copy global -> shared; bytes = 0 zero-fills the destination
__device__ __forceinline__ void as_cp16(double* dst, const double* src, int bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

// Tile geometry as template parameters: TM rows per tile, thread tile TR rows x TC columns (TR * TC = TM / 8), MINB CTAs per SM.
//   <128, 4, 4, 2> is the production geometry: two CTAs per SM overlap the sequential phases of the factorisation (diagonal
//   block by one warp, panel solve by one thread per row) with the other CTA's update.  <256, 2, 16, 1> (3.3x less shared-memory
//   traffic per FMA, but one CTA per SM) was measured slower even on the heavy columns alone (config 3a sweep 8.0 s vs 5.9 s):
//   the kernel is bound by those serial phases and by latency, not by shared-memory or HBM bandwidth.
template <int TM, int TR, int TC, int MINB>
__global__ void __launch_bounds__(AS_THREADS, MINB) kf_as_chol_kernel(const AsArgs a) {
    constexpr int TLD = TM + 1;
    constexpr int RT = TM / TR;            // row threads; column groups = AS_THREADS / RT = AS_NB / TC
    static_assert(TR * TC * AS_THREADS == TM * AS_NB && (AS_THREADS / RT) * TC == AS_NB && TC % 2 == 0, "tile geometry");
    extern __shared__ __align__(16) double as_smem[];
    double* sA = as_smem;                              // [AS_NST][AS_KC][TM]  k-chunks of the row-tile operand (cp.async ring)
    double* sB = sA + AS_NST * AS_KC * TM;          // [AS_NST][AS_KC][AS_NB]  k-chunks of the block-row operand: sB[k][c]
    double* sD = sB + AS_NST * AS_KC * AS_NB;          // [AS_NB][AS_NB+1] diagonal block factor: sD[r][c]
    double* sInv = sD + AS_NB * (AS_NB + 1);           // [AS_NB] 1 / diag
    int* sDead = reinterpret_cast<int*>(sInv + AS_NB); // [AS_NB]
    int* sIdx = sDead + AS_NB;                         // [P]
    double* sT = sA;                                   // [AS_NB][TLD]  updated tile, column-major (aliases the ring: idle after the k loop)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col = a.cols[blockIdx.x];
    const int n = a.cnt[col];
    // this column's a_j, b_j are rewritten on the new support at the end: clear what the old support left behind (columns whose
    // support did not change are not launched at all and keep their a_j, b_j, s'a, s'b)
    if (a.mode != 1)
        for (int i = tid; i < a.P; i += AS_THREADS) {
            a.Aout[i + (long long)col * a.ld] = 0.0;
            a.Bout[i + (long long)col * a.ld] = 0.0;
        }
    if (n == 0) {
        if (tid == 0 && a.mode != 1) { a.col_num[col] = 0.0; a.col_den[col] = 0.0; a.col_dead[col] = 0; }
        return;
    }
    if (a.mode == 1 && (a.J0 >= n || a.J0 + (blockIdx.y + a.tile_off) * TM >= n + 2)) return;      // nothing of this tile exists
    const int ldl = as_ldl(n);
    double* __restrict__ L = a.ws + a.ws_off[blockIdx.x];
    const int* gidx = a.idx + (long long)col * a.P;
    for (int i = tid; i < n; i += AS_THREADS) sIdx[i] = gidx[i];
    __syncthreads();
    const double* Ccol = a.C + (long long)col * a.ld;
    const double* Scol = a.SG + (long long)col * a.ld;
    const int nrows = n + 2;      // rows n, n+1: the right-hand sides c_S and s_S
    int ndead = 0;

    const int tr = tid % RT, tc = tid / RT;   // thread owns tile rows tr + RT i (i < TR) and tile columns TC tc + j (j < TC)
    int* gdead = a.dead ? a.dead + (long long)blockIdx.x * a.P : nullptr;
    const int jb_first = a.mode == 1 ? a.J0 : 0, jb_end = a.mode == 0 ? n : (a.mode == 1 ? min(n, a.J0 + 1) : 0);
    for (int J0 = jb_first; J0 < jb_end; J0 += AS_NB) {
        const int w = min(AS_NB, n - J0);
        const int r_first = a.mode == 1 ? J0 + (blockIdx.y + a.tile_off) * TM : J0;
        const int r_end = a.mode == 1 ? min(nrows, r_first + 1) : nrows;
        for (int R0 = r_first; R0 < r_end; R0 += TM) {
            double acc[TR][TC];
#pragma unroll
            for (int i = 0; i < TR; ++i)
#pragma unroll
                for (int j = 0; j < TC; ++j) acc[i][j] = 0.0;
            // ---- acc = L[R0.., 0:J0] * L[J0.., 0:J0]': k-chunks of AS_KC through an AS_NST-deep cp.async ring, so that the
            // factor data (HBM: the factors of the ~300 columns in flight do not fit L2) is requested 3 chunks ahead.
            // Pairs of consecutive rows (16 bytes; ldl, R0, J0 are even and L is 16-byte aligned)
            auto stage = [&](int k0, int b) {
                double* dA = sA + b * AS_KC * TM;
                double* dB = sB + b * AS_KC * AS_NB;
#pragma unroll
                for (int i = 0; i < (AS_KC * TM / 2) / AS_THREADS; ++i) {
                    const int e = tid + i * AS_THREADS;     // pair index: kk * 64 + r / 2
                    const int r = (e & (TM / 2 - 1)) * 2, kk = e / (TM / 2);
                    const int gr = R0 + r;
                    as_cp16(dA + kk * TM + r, L + (gr < nrows ? gr : 0) + (long long)(k0 + kk) * ldl, gr < nrows ? 16 : 0);
                }
                {
                    const int e = tid;                      // pair index: kk * 16 + c / 2   (AS_KC * AS_NB / 2 = 256 pairs)
                    const int c = (e & (AS_NB / 2 - 1)) * 2, kk = e / (AS_NB / 2);
                    as_cp16(dB + kk * AS_NB + c, L + (J0 + (c < w ? c : 0)) + (long long)(k0 + kk) * ldl, c < w ? 16 : 0);
                }
            };
            const int nchunk = J0 / AS_KC;
#pragma unroll
            for (int p = 0; p < AS_NST - 1; ++p) {
                if (p < nchunk) stage(p * AS_KC, p);
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            for (int q = 0; q < nchunk; ++q) {
                asm volatile("cp.async.wait_group %0;" ::"n"(AS_NST - 2) : "memory");
                __syncthreads();                            // chunk q has landed; everyone is done with chunk q - 1
                if (q + AS_NST - 1 < nchunk) stage((q + AS_NST - 1) * AS_KC, (q + AS_NST - 1) % AS_NST);
                asm volatile("cp.async.commit_group;" ::: "memory");
                const double* cA = sA + (q % AS_NST) * AS_KC * TM + tr;
                const double* cB = sB + (q % AS_NST) * AS_KC * AS_NB + TC * tc;
#pragma unroll 4
                for (int kk = 0; kk < AS_KC; ++kk) {
                    double av[TR], bv[TC];
#pragma unroll
                    for (int i = 0; i < TR; ++i) av[i] = cA[kk * TM + RT * i];
#pragma unroll
                    for (int j = 0; j < TC / 2; ++j) {
                        const double2 t2 = *reinterpret_cast<const double2*>(cB + kk * AS_NB + 2 * j);
                        bv[2 * j] = t2.x; bv[2 * j + 1] = t2.y;
                    }
#pragma unroll
                    for (int i = 0; i < TR; ++i)
#pragma unroll
                        for (int j = 0; j < TC; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
                }
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                // the ring is idle: sT may overwrite it
            // ---- tile = A[R0.., J0..] - acc, A gathered from G (rows < n), c_S (row n), s_S (row n+1)
#pragma unroll
            for (int j = 0; j < TC; ++j) {
                const int c = TC * tc + j;
                const int gc = c < w ? sIdx[J0 + c] : 0;
#pragma unroll
                for (int i = 0; i < TR; ++i) {
                    const int r = tr + RT * i, gr = R0 + r;
                    double v = 0.0;
                    if (c < w && gr < nrows && gr >= J0 + c) {
                        if (gr < n) v = a.G[sIdx[gr] + (long long)gc * a.ldg];
                        else v = (gr == n) ? Ccol[gc] : Scol[gc];
                        v -= acc[i][j];
                    }
                    sT[c * TLD + r] = v;
                }
            }
            __syncthreads();
            if (R0 == J0) {
                // ---- factor the w x w diagonal block (tile rows 0..w-1) with warp 0; lane = row
                if (warp == 0) {
                    // lane r keeps row r of the block in registers (x[k] = T[r][k], k <= r); column c of the factor is broadcast
                    // lane to lane with shuffles — no shared-memory round trips inside the 32 dependent steps
                    double x[AS_NB];
#pragma unroll
                    for (int k = 0; k < AS_NB; ++k) x[k] = (lane < w && k <= lane) ? sT[k * TLD + lane] : 0.0;
                    const double g0 = lane < w ? a.G[sIdx[J0 + lane] + (long long)sIdx[J0 + lane] * a.ldg] : 1.0;
#pragma unroll
                    for (int c = 0; c < AS_NB; ++c) {
                        if (c < w) {                                   // uniform
                            const double p = __shfl_sync(0xffffffffu, x[c], c);
                            const double gc = __shfl_sync(0xffffffffu, g0, c);
                            const bool dead = !(p > 2e-15 * gc);
                            const double d = dead ? 1.0 : sqrt(p);
                            const double inv = 1.0 / d;
                            const double lrc = (lane > c && !dead) ? x[c] * inv : 0.0;     // L[lane][c]
#pragma unroll
                            for (int k = c + 1; k < AS_NB; ++k) {
                                const double lk = __shfl_sync(0xffffffffu, lrc, k);        // L[k][c]
                                x[k] = fma(-lrc, lk, x[k]);                               // only k <= lane is ever read
                            }
                            x[c] = (lane == c) ? d : lrc;
                            if (lane == c) { sInv[c] = inv; sDead[c] = dead ? 1 : 0; }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < AS_NB; ++c)
                        if (c < w && lane < w) sD[lane * (AS_NB + 1) + c] = lane >= c ? x[c] : 0.0;
                }
                __syncthreads();
                if (a.mode == 1 && tid < w) gdead[J0 + tid] = sDead[tid];
            } else if (a.mode == 1) {
                // a later tile of the block column: the diagonal factor was written by the tile-0 launch
                for (int e = tid; e < AS_NB * AS_NB; e += AS_THREADS) {
                    const int r = e % AS_NB, c = e / AS_NB;
                    sD[r * (AS_NB + 1) + c] = (r < w && c <= r) ? L[(J0 + r) + (long long)(J0 + c) * ldl] : 0.0;
                }
                if (tid < AS_NB) sDead[tid] = tid < w ? gdead[J0 + tid] : 0;
                __syncthreads();
                if (tid < AS_NB) sInv[tid] = tid < w ? 1.0 / sD[tid * (AS_NB + 1) + tid] : 0.0;
                __syncthreads();
            }
            // ---- rows below the diagonal block: X = T * L_D^-T (one thread per row), write the block column of L
            if (tid < TM) {
                const int r = tid, gr = R0 + r;
                if (gr < nrows) {
                    if (gr < J0 + w) {               // a row of the diagonal block
                        for (int c = 0; c <= gr - J0; ++c) L[gr + (long long)(J0 + c) * ldl] = sD[(gr - J0) * (AS_NB + 1) + c];
                    } else {
                        double x[AS_NB];
#pragma unroll
                        for (int c = 0; c < AS_NB; ++c) x[c] = c < w ? sT[c * TLD + r] : 0.0;
#pragma unroll
                        for (int c = 0; c < AS_NB; ++c) {
                            if (c < w) {
                                double s = x[c];
#pragma unroll
                                for (int k = 0; k < c; ++k) s = fma(-x[k], sD[c * (AS_NB + 1) + k], s);
                                x[c] = sDead[c] ? 0.0 : s * sInv[c];
                                L[gr + (long long)(J0 + c) * ldl] = x[c];
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (tid == 0 && (a.mode == 0 || r_first == J0))
            for (int c = 0; c < w; ++c) ndead += sDead[c];
    }
    if (a.mode == 1) {          // only the tile-0 CTA counted: one writer per column and launch
        if (tid == 0 && blockIdx.y + a.tile_off == 0) a.col_dead[col] = (a.J0 == 0 ? 0 : a.col_dead[col]) + ndead;
        return;
    }

    // ---- back-substitution L' x = y for both right-hand sides (y = rows n, n+1 of L); x lives in shared memory
    double* xs = as_smem;            // [2][n] (n <= 4096), aliases the ring
    for (int i = tid; i < n; i += AS_THREADS) {
        xs[i] = L[n + (long long)i * ldl];
        xs[n + i] = L[n + 1 + (long long)i * ldl];
    }
    __syncthreads();
    const int nblk = (n + AS_NB - 1) / AS_NB;
    for (int jb = nblk - 1; jb >= 0; --jb) {
        const int J0 = jb * AS_NB, w = min(AS_NB, n - J0);
        // contributions of the rows already solved
        for (int c = warp; c < w; c += AS_THREADS / 32) {
            const double* Lc = L + (long long)(J0 + c) * ldl;
            double s0 = 0.0, s1 = 0.0;
            for (int r = J0 + w + lane; r < n; r += 32) {
                const double l = Lc[r];
                s0 = fma(l, xs[r], s0);
                s1 = fma(l, xs[n + r], s1);
            }
            for (int off = 16; off > 0; off >>= 1) {
                s0 += __shfl_down_sync(0xffffffffu, s0, off);
                s1 += __shfl_down_sync(0xffffffffu, s1, off);
            }
            if (lane == 0) { xs[J0 + c] -= s0; xs[n + J0 + c] -= s1; }
        }
        for (int e = tid; e < w * w; e += AS_THREADS) {
            const int r = e % w, c = e / w;
            sD[r * (AS_NB + 1) + c] = r >= c ? L[(J0 + r) + (long long)(J0 + c) * ldl] : 0.0;
        }
        __syncthreads();
        if (warp < 2) {                  // warp q solves right-hand side q with L_D' (upper triangular)
            double x = lane < w ? xs[warp * n + J0 + lane] : 0.0;
            for (int c = w - 1; c >= 0; --c) {
                if (lane == c) x = x / sD[c * (AS_NB + 1) + c];
                const double xc = __shfl_sync(0xffffffffu, x, c);
                if (lane < c) x = fma(-sD[c * (AS_NB + 1) + lane], xc, x);
            }
            if (lane < w) xs[warp * n + J0 + lane] = x;
        }
        __syncthreads();
    }
    // ---- scatter a_j, b_j and the column sums s'a, s'b
    double num = 0.0, den = 0.0;
    for (int i = tid; i < n; i += AS_THREADS) {
        const int gi = sIdx[i];
        const double s = Scol[gi], xa = xs[i], xb = xs[n + i];
        a.Aout[gi + (long long)col * a.ld] = xa;
        a.Bout[gi + (long long)col * a.ld] = xb;
        num = fma(s, xa, num);
        den = fma(s, xb, den);
    }
    __shared__ double red[2][AS_THREADS / 32];
    for (int off = 16; off > 0; off >>= 1) {
        num += __shfl_down_sync(0xffffffffu, num, off);
        den += __shfl_down_sync(0xffffffffu, den, off);
    }
    if (lane == 0) { red[0][warp] = num; red[1][warp] = den; }
    __syncthreads();
    if (tid == 0) {
        double sn = 0.0, sd = 0.0;
        for (int q = 0; q < AS_THREADS / 32; ++q) { sn += red[0][q]; sd += red[1][q]; }
        a.col_num[col] = sn;
        a.col_den[col] = sd;
        if (a.mode == 0) a.col_dead[col] = ndead;
    }
}

// one CTA per column: ascending list of the support {i : SG(i, col) != 0} and its size
__global__ void __launch_bounds__(256) kf_as_index_kernel(const double* SG, long long ld, int P, AsCols cs, int* idx, int* cnt) {
    __shared__ int wsum[8];
    __shared__ int base;
    const int col = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base = 0;
    __syncthreads();
    if (!cs.on(col)) {
        if (tid == 0) cnt[col] = 0;
        return;
    }
    const double* s = SG + (long long)col * ld;
    int* out = idx + (long long)col * P;
    for (int i0 = 0; i0 < P; i0 += 256) {
        const int i = i0 + tid;
        const bool on = i < P && s[i] != 0.0;
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int q = 0; q < warp; ++q) off += wsum[q];
        if (on) out[off + __popc(m & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int q = 0; q < 8; ++q) t += wsum[q];
            base += t;
        }
        __syncthreads();
    }
    if (tid == 0) cnt[col] = base;
}

// single CTA: scal = [sum_j s'a_j, sum_j s'b_j, skipped pivots, support size] over this rank's columns (fixed order: deterministic)
__global__ void __launch_bounds__(256) kf_as_sums_kernel(const double* col_num, const double* col_den, const int* col_dead, const int* cnt,
                                                         int P, AsCols cs, double* scal) {
    __shared__ double r0[256], r1[256], r2[256], r3[256];
    double n = 0.0, d = 0.0, dd = 0.0, nz = 0.0;
    for (int i = threadIdx.x; i < P; i += 256) {
        if (!cs.on(i)) continue;
        n += col_num[i]; d += col_den[i]; dd += col_dead[i]; nz += cnt[i];
    }
    r0[threadIdx.x] = n; r1[threadIdx.x] = d; r2[threadIdx.x] = dd; r3[threadIdx.x] = nz;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            r0[threadIdx.x] += r0[threadIdx.x + s];
            r1[threadIdx.x] += r1[threadIdx.x + s];
            r2[threadIdx.x] += r2[threadIdx.x + s];
            r3[threadIdx.x] += r3[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { scal[0] = r0[0]; scal[1] = r1[0]; scal[2] = r2[0]; scal[3] = r3[0]; }   // [3]: support size
}

// Pattern changes that trial multipliers would cause: K(lam) = A - lam B and grad(lam) = G A - lam G B - C are both
// affine in lam, so AS_NL trials cost ONE element-wise pass and no factorisation.  counts[l] = leaving + entering for lams[l]
struct AsLams { double v[AS_NL]; };
__global__ void __launch_bounds__(256) kf_as_count_kernel(const double* A, const double* B, const double* SG, const double* GA, const double* GB,
                                                          const double* C, long long ld, int P, AsCols cs, AsLams lams, int nl, double rel,
                                                          unsigned long long* counts) {
    __shared__ unsigned int sc[AS_NL];
    if (threadIdx.x < AS_NL) sc[threadIdx.x] = 0;
    __syncthreads();
    const int ncol = cs.hi - cs.lo;
    const long long n = (long long)P * ncol;
    int cnt[AS_NL];
#pragma unroll
    for (int l = 0; l < AS_NL; ++l) cnt[l] = 0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % P), c = cs.lo + (int)(e / P);
        if (c >= cs.skip0 && c < cs.skip1) continue;
        const long long o = i + (long long)c * ld;
        const double s = SG[o];
        if (s != 0.0) {
            const double a = A[o], b = B[o];
#pragma unroll
            for (int l = 0; l < AS_NL; ++l)
                if (l < nl && !(fma(-lams.v[l], b, a) * s > 0.0)) ++cnt[l];
        } else {
            const double ga = GA[o] - C[o], gb = GB[o];
#pragma unroll
            for (int l = 0; l < AS_NL; ++l)
                if (l < nl && fabs(fma(-lams.v[l], gb, ga)) > lams.v[l] * (1.0 + rel)) ++cnt[l];
        }
    }
#pragma unroll
    for (int l = 0; l < AS_NL; ++l) {
        int v = cnt[l];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sc[l], (unsigned)v);
    }
    __syncthreads();
    if (threadIdx.x < nl && sc[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)sc[threadIdx.x]);
}

// K = A - lam B on the support; entries whose sign flipped leave (K = 0), dual violators enter with sign -sign(grad)
__global__ void kf_as_apply_kernel(const double* A, const double* B, double* SG, const double* GA, const double* GB, const double* C, double* K,
                                   long long ld, int P, AsCols cs, double lam, double rel, int* chg) {
    const long long n = (long long)P * (cs.hi - cs.lo);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % P), c = cs.lo + (int)(e / P);
        if (c >= cs.skip0 && c < cs.skip1) continue;
        const long long o = i + (long long)c * ld;
        const double s = SG[o];
        double k = 0.0;
        if (s != 0.0) {
            k = fma(-lam, B[o], A[o]);
            if (!(k * s > 0.0)) { k = 0.0; SG[o] = 0.0; chg[c] = 1; }
        } else {
            const double g = fma(-lam, GB[o], GA[o]) - C[o];
            if (fabs(g) > lam * (1.0 + rel)) { SG[o] = g > 0.0 ? -1.0 : 1.0; chg[c] = 1; }
        }
        K[o] = k;
    }
}

// Diagnostic (option as_diag): how much of every column's factor would survive a step if the support were kept in order of
// entry ("age") instead of ascending index — leaving entries invalidate the factor from their position on, entering ones
// are appended.  birth[i][c] = step at which entry (i, c) joined the support (0: not in it).  stats += [flops of a full
// refactorisation, flops with reuse, entering, leaving, columns whose support changed, columns]
__global__ void __launch_bounds__(256) kf_as_diag_kernel(const double* A, const double* B, const double* SG, const double* GA, const double* GB,
                                                         const double* C, long long ld, int P, AsCols cs, double lam, double rel, int stepno,
                                                         int* birth, double* stats) {
    const int c = cs.lo + blockIdx.x;
    if (!cs.on(c)) return;
    __shared__ int s_min, s_keep, s_n, s_in, s_out;
    if (threadIdx.x == 0) { s_min = 0x7fffffff; s_keep = 0; s_n = 0; s_in = 0; s_out = 0; }
    __syncthreads();
    int* bc = birth + (long long)c * P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const long long o = i + (long long)c * ld;
        const double sg = SG[o];
        if (sg != 0.0) {
            atomicAdd(&s_n, 1);
            if (bc[i] == 0) bc[i] = 1;      // cold-start entries
            const double k = fma(-lam, B[o], A[o]);
            if (!(k * sg > 0.0)) { atomicMin(&s_min, bc[i]); atomicAdd(&s_out, 1); }
        } else {
            const double g = fma(-lam, GB[o], GA[o]) - C[o];
            if (fabs(g) > lam * (1.0 + rel)) atomicAdd(&s_in, 1);
        }
    }
    __syncthreads();
    const int bmin = s_min;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const long long o = i + (long long)c * ld;
        const double sg = SG[o];
        if (sg != 0.0) {
            const double k = fma(-lam, B[o], A[o]);
            if (!(k * sg > 0.0)) bc[i] = 0;
            else if (bc[i] < bmin) atomicAdd(&s_keep, 1);
        } else {
            const double g = fma(-lam, GB[o], GA[o]) - C[o];
            if (fabs(g) > lam * (1.0 + rel)) bc[i] = stepno;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double n1 = (double)(s_n - s_out + s_in), keep = (double)s_keep;
        atomicAdd(stats + 0, n1 * n1 * n1 / 3.0);
        atomicAdd(stats + 1, (s_in + s_out) ? (n1 - keep) * keep * keep + (n1 * n1 * n1 - keep * keep * keep) / 3.0 : 0.0);
        atomicAdd(stats + 2, (double)s_in);
        atomicAdd(stats + 3, (double)s_out);
        atomicAdd(stats + 4, (s_in + s_out) ? 1.0 : 0.0);
        atomicAdd(stats + 5, 1.0);
    }
}

// cold start: SG = sign(C) where |C| >= thr (free columns), 0 elsewhere;  warm start: SG = sign(K)
__global__ void kf_as_init_kernel(const double* src, double thr, double* SG, long long ld, int P, int Pp, AsCols cs) {
    const long long n = (long long)Pp * Pp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % Pp), c = (int)(e / Pp);
        double s = 0.0;
        if (i < P && c < P && cs.on(c)) {
            const double v = src[i + (long long)c * ld];
            if (fabs(v) >= thr && v != 0.0) s = v > 0.0 ? 1.0 : -1.0;
        }
        SG[i + (long long)c * ld] = s;
    }
}

__global__ void kf_as_absmax_kernel(const double* C, long long ld, int P, AsCols cs, double* out) {
    __shared__ double red[32];
    double m = 0;
    const long long n = (long long)P * P;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
        const int i = (int)(e % P), c = (int)(e / P);
        if (!cs.on(c)) continue;
        m = fmax(m, fabs(C[i + (long long)c * ld]));
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

}  // namespace

// Solve the nb ACTIVE budgets t[0..nb) (any order; traversed ascending, each warm-started from the previous support).
// G, C: Pp-strided P x P.  K_all: nb matrices (stride Pp*Pp); the pinned delay columns [fix_c0, fix_c1) must already
// hold their pattern in every K_b and t[] is the budget left for the free columns.
// Column partition (kf_set_qp_partition): this rank factors and updates only the columns [qp_lo, qp_hi) of K — the
// columns couple through the multiplier and the step control alone, i.e. through a handful of scalars per step that are
// summed over the ranks by the caller's all-reduce hook; every rank then takes identical decisions.
int kf_solve_l1ball_as(kf_ctx* ctx, int P, int Pp, const double* G, const double* C, int nb, const double* t, int fix_c0, int fix_c1,
                       int max_iter, double* K_all, KfQpResult* res, cudaStream_t st) {
    if (nb <= 0) return KF_OK;
    if (P > 4096) {
        ctx->err = "active-set QP solver: P > 4096 is not supported";
        return KF_EUNSUPPORTED;
    }
    const bool split = ctx->qp_hi > ctx->qp_lo;
    AsCols cs{split ? std::max(ctx->qp_lo, 0) : 0, split ? std::min(ctx->qp_hi, P) : P, fix_c0, fix_c1};
    if (split && fix_c1 > fix_c0) {
        ctx->err = "active-set QP solver: a column partition cannot be combined with pinned delay columns";
        return KF_EUNSUPPORTED;
    }
    // Two ways to sum the step scalars over the ranks: the library's own NCCL communicator (kf_comm_init_rank /
    // kf_create_multi) reduces the small DEVICE buffers in stream order, before they are read back — no host staging, no
    // extra synchronisation; a caller-supplied hook (kf_set_qp_partition) reduces the host copies.
    const bool dev_reduce = split && !ctx->qp_allreduce && ctx->comm && ctx->nranks > 1;
    auto reduce = [&](double* v, int n, int op) -> int {       // op 0 = sum, 1 = max over the ranks
        if (!split || !ctx->qp_allreduce) return KF_OK;
        if (ctx->qp_allreduce(ctx->qp_user, v, n, op)) {
            ctx->err = "active-set QP solver: the all-reduce hook failed";
            return KF_EINVAL;
        }
        return KF_OK;
    };
    const int ncol = cs.hi - cs.lo;
    const long long ld = Pp;
    const size_t mat = (size_t)Pp * Pp;
    // matrices: SG | A | B | GA | GB ; ints: idx[P*P] | cnt[P] | cols[P] | dead[P] ; doubles: num[P] | den[P] | scal[8] ; offsets[P]
    KF_CUDA(ctx, ctx->d_as_mat.ensure(5 * mat * sizeof(double)));
    double* SG = ctx->d_as_mat.as<double>();
    double* Am = SG + mat;
    double* Bm = Am + mat;
    double* GA = Bm + mat;      // G A
    double* GB = GA + mat;      // G B
    const size_t n_int = 2ull * P * P + 4ull * P + 8;
    const size_t n_dbl = 2ull * P + 8;
    KF_CUDA(ctx, ctx->d_as_aux.ensure(n_dbl * sizeof(double) + (size_t)(P + AS_NL + 2) * sizeof(long long) + n_int * sizeof(int) + 64));
    double* d_num = ctx->d_as_aux.as<double>();
    double* d_den = d_num + P;
    double* d_scal = d_den + P;
    long long* d_off = reinterpret_cast<long long*>(d_scal + 8);
    unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(d_off + P);
    int* d_idx = reinterpret_cast<int*>(d_counts + AS_NL + 2);
    int* d_cnt = d_idx + (size_t)P * P;
    int* d_cols = d_cnt + P;
    int* d_dead = d_cols + P;
    int* d_chg = d_dead + P;        // [column]: support changed since its last factorisation
    int* d_deadflags = d_chg + P + 8;   // [launch slot][P]: skipped pivots (level-synchronous mode)

    // workspace for the factors: bounded (option as_ws_gb, default 8 GB), columns are processed in chunks that fit
    size_t free_b = 0, total_b = 0;
    KF_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    const size_t one_col = (size_t)(P + 3) * P * sizeof(double);
    size_t ws_bytes = std::min<size_t>((size_t)(ctx->opt_as_ws_gb * 1073741824.0), (size_t)std::max(ncol, 1) * one_col);
    ws_bytes = std::max(ws_bytes, one_col);
    if (ctx->d_as_ws.bytes < ws_bytes) {
        if (ws_bytes > ctx->d_as_ws.bytes + free_b) ws_bytes = std::max(one_col, (size_t)(0.8 * (double)(ctx->d_as_ws.bytes + free_b)));
        KF_CUDA(ctx, ctx->d_as_ws.ensure(ws_bytes));
    }
    const size_t ws_doubles = ctx->d_as_ws.bytes / sizeof(double);

    auto smem_of = [&](int tm) {
        return (size_t)(AS_NST * AS_KC * (tm + AS_NB) + AS_NB * (AS_NB + 1) + AS_NB) * sizeof(double) + (size_t)(AS_NB + P) * sizeof(int) + 16;
    };
    const size_t smem = smem_of(128);
    KF_CUDA(ctx, kf_ensure_smem(ctx, kf_as_chol_kernel<128, 4, 4, 2>, smem));
    const int egrid = ctx->sm_count * 4;
    const int iter_cap = max_iter > 0 ? max_iter : 200;
    const double rel = 1e-10;

    std::vector<int> order(nb);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return t[x] < t[y]; });

    KfBuf diag_birth;                     // option as_diag only
    int diag_step = 0;
    double seg[4] = {0, 0, 0, 0};         // as_diag: host wall time of the step segments (index | factor + GEMM | counts | rest)
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    std::vector<int> h_cnt(P), h_cols(P), h_chg(P);
    long long cols_factored = 0;
    std::vector<long long> h_off(P);
    unsigned long long h_counts[AS_NL];
    double lam_prev = 0.0, lam_now = 0.0;
    double frac = ctx->opt_as_frac;       // bound on the pattern change per step; halved when a budget fails to contract
    bool clamped = false;
    long long changed = 0;

    // counts[l] = entries that would leave or enter at lams[l], summed over the ranks
    auto count = [&](const double* lams, int nl, double* out) -> int {
        AsLams L{};
        for (int l = 0; l < nl; ++l) L.v[l] = lams[l];
        KF_CUDA(ctx, cudaMemsetAsync(d_counts, 0, AS_NL * sizeof(unsigned long long), st));
        kf_as_count_kernel<<<egrid, 256, 0, st>>>(Am, Bm, SG, GA, GB, C, ld, P, cs, L, nl, rel, d_counts);
        KF_CUDA(ctx, cudaGetLastError());
        if (dev_reduce) KF_TRY(kf_comm_allreduce_u64(ctx, d_counts, AS_NL, st));
        KF_CUDA(ctx, cudaMemcpyAsync(h_counts, d_counts, AS_NL * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->launches += 1;
        for (int l = 0; l < nl; ++l) out[l] = (double)h_counts[l];
        return reduce(out, nl, 0);
    };

    // one exact step on the current pattern SG for budget tb; K -> Kout, pattern updated
    auto step = [&](double tb, double* Kout) -> int {
        const auto ts0 = tnow();
        kf_as_index_kernel<<<P, 256, 0, st>>>(SG, ld, P, cs, d_idx, d_cnt);
        KF_CUDA(ctx, cudaMemcpyAsync(h_cnt.data(), d_cnt, sizeof(int) * P, cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaMemcpyAsync(h_chg.data(), d_chg, sizeof(int) * P, cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaMemsetAsync(d_chg, 0, sizeof(int) * P, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        // only the columns whose support changed are factored again: a_j = G_SS^-1 c_S and b_j = G_SS^-1 s_S depend on the pattern
        // alone (config 3a: about half of all column-steps are unchanged)
        const auto ts1 = tnow();
        int nown = 0;
        for (int j = cs.lo; j < cs.hi; ++j)
            if (cs.on(j) && (h_chg[j] || !ctx->opt_as_skip)) h_cols[nown++] = j;
        cols_factored += nown;
        std::stable_sort(h_cols.begin(), h_cols.begin() + nown, [&](int x, int y) { return h_cnt[x] > h_cnt[y]; });   // heavy columns first
        AsArgs a{};
        a.G = G; a.ldg = ld; a.C = C; a.SG = SG; a.Aout = Am; a.Bout = Bm; a.ld = ld;
        a.idx = d_idx; a.cnt = d_cnt; a.ws = ctx->d_as_ws.as<double>();
        a.col_num = d_num; a.col_den = d_den; a.col_dead = d_dead; a.P = P;
        // chunks of columns whose factors fit the workspace together; one upload, kernels back to back on the stream
        std::vector<int> chunk_end;
        {
            size_t used = 0;
            for (int c = 0; c < nown; ++c) {
                const int n = h_cnt[h_cols[c]];
                const size_t need = (size_t)((n + 3) & ~1) * (size_t)std::max(n, 1);
                if (used > 0 && used + need > ws_doubles) { chunk_end.push_back(c); used = 0; }
                h_off[c] = (long long)used;
                used += need;
            }
            chunk_end.push_back(nown);
        }
        if (nown) {
            KF_CUDA(ctx, cudaMemcpyAsync(d_cols, h_cols.data(), sizeof(int) * nown, cudaMemcpyHostToDevice, st));
            KF_CUDA(ctx, cudaMemcpyAsync(d_off, h_off.data(), sizeof(long long) * nown, cudaMemcpyHostToDevice, st));
        }
        // Few columns for the CTA slots of this GPU (a rank of a column-split sweep): one CTA per column would make the step as long
        // as ONE heavy column's ~30 dependent block columns x its row tiles.  Level-synchronous instead: per block column one launch
        // for the diagonal tiles and one for all the other row tiles of all columns, then the back-substitutions.
        const bool level = nown > 0 && chunk_end.size() == 1 &&
                           (ctx->opt_as_level == 1 || (ctx->opt_as_level == 0 && nown <= 2 * ctx->sm_count && h_cnt[h_cols[0]] >= 256));
        a.dead = d_deadflags;
        if (level) {
            a.cols = d_cols; a.ws_off = d_off;
            const int nmax = h_cnt[h_cols[0]];
            int active = nown;
            for (int J0 = 0; J0 < nmax; J0 += AS_NB) {
                while (active > 0 && h_cnt[h_cols[active - 1]] <= J0) --active;      // heavy first: the live columns are a prefix
                a.mode = 1; a.J0 = J0; a.tile_off = 0;
                kf_as_chol_kernel<128, 4, 4, 2><<<dim3(active, 1), AS_THREADS, smem, st>>>(a);
                const int ntiles = (nmax + 2 - J0 + 127) / 128;
                if (ntiles > 1) {
                    a.tile_off = 1;
                    kf_as_chol_kernel<128, 4, 4, 2><<<dim3(active, ntiles - 1), AS_THREADS, smem, st>>>(a);
                }
                ctx->launches += 2;
            }
            a.mode = 2; a.J0 = 0; a.tile_off = 0;
            kf_as_chol_kernel<128, 4, 4, 2><<<nown, AS_THREADS, smem, st>>>(a);
            KF_CUDA(ctx, cudaGetLastError());
            ctx->launches += 1;
        } else {
            a.mode = 0;
            int c0 = 0;
            for (int c1 : chunk_end) {
                if (c1 > c0) {
                    a.cols = d_cols + c0; a.ws_off = d_off + c0;
                    kf_as_chol_kernel<128, 4, 4, 2><<<c1 - c0, AS_THREADS, smem, st>>>(a);
                    KF_CUDA(ctx, cudaGetLastError());
                    ctx->launches += 1;
                }
                c0 = c1;
            }
        }
        kf_as_sums_kernel<<<1, 256, 0, st>>>(d_num, d_den, d_dead, d_cnt, P, cs, d_scal);
        if (dev_reduce) KF_TRY(kf_comm_allreduce(ctx, d_scal, 4, 0, st));   // s'a, s'b, skipped pivots, support size over the ranks
        double sums[4] = {0, 0, 0, 0};
        KF_CUDA(ctx, cudaMemcpyAsync(sums, d_scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (ncol > 0)
            for (int q = 0; q < 2; ++q) {        // GA = G A, GB = G B on this rank's columns  (G symmetric: row i of G = column i)
                KfGemmGrid g{};
                g.A = G; g.lda = ld; g.B = (q ? Bm : Am) + (size_t)cs.lo * ld; g.ldb = ld; g.out = (q ? GB : GA) + (size_t)cs.lo * ld;
                g.ldm = 1; g.ldn = ld;
                g.m = P; g.n = ncol; g.k0 = 0; g.k1 = Pp; g.alpha = 1.0; g.accumulate = 0; g.lower_only = 0;
                KF_TRY(kf_launch_gemm_grid(ctx, g, st));
            }
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        const auto ts2 = tnow();
        KF_TRY(reduce(sums, 4, 0));                 // hook path: s'a, s'b, skipped pivots, support size summed over the ranks
        const double nnz = sums[3];
        const double limit = std::max(frac * nnz, (double)P);
        // The multiplier that meets the budget on the current pattern, lam_b = (sum s'a - t) / sum s'b, under two step
        // controls.  (1) lam at most halves per step: a budget the current support cannot use up would give lam = 0,
        // every off-support entry would enter at once and the iteration never recovers (observed).  (2) The step towards
        // that goal is shortened until the pattern change stays below `limit`: a step that swaps a large part of the
        // support at once is no contraction any more (observed from ~40 % density on).
        const double lam_b = sums[1] > 0.0 ? std::max((sums[0] - tb) / sums[1], 0.0) : 0.0;
        const double lam_goal = std::max(lam_b, 0.5 * lam_prev);
        clamped = lam_goal > lam_b;
        double lam = lam_goal, cnts[AS_NL], lams[AS_NL];
        if (!(lam_prev > 0.0) || lam_goal == lam_prev) {
            KF_TRY(count(&lam_goal, 1, cnts));
            changed = (long long)cnts[0];
        } else {
            // theta in (0, 1]: lam = lam_prev + theta (lam_goal - lam_prev); a grid of AS_NL trials per pass, theta = 1 included
            double lo = 0.0, hi = 1.0;
            for (int pass = 0; pass < 2; ++pass) {
                const double w = (hi - lo) / AS_NL;
                for (int l = 0; l < AS_NL; ++l) lams[l] = lam_prev + (lo + w * (l + 1)) * (lam_goal - lam_prev);
                if (pass == 0) lams[AS_NL - 1] = lam_goal;
                KF_TRY(count(lams, AS_NL, cnts));
                if (pass == 0 && cnts[AS_NL - 1] <= limit) {      // the full step is fine
                    changed = (long long)cnts[AS_NL - 1];
                    lo = 1.0;
                    break;
                }
                int best = -1;
                for (int l = 0; l < AS_NL; ++l)
                    if (cnts[l] <= limit) best = l; else break;
                if (best >= 0) changed = (long long)cnts[best];
                lo = lo + w * (best + 1);
                hi = lo + w;
                clamped = true;
            }
            lam = lo >= 1.0 ? lam_goal : lam_prev + lo * (lam_goal - lam_prev);
            if (lo == 0.0) {                      // even the smallest step swaps too much: settle the pattern at lam_prev first
                KF_TRY(count(&lam, 1, cnts));
                changed = (long long)cnts[0];
            }
        }
        if (ctx->opt_as_diag) {
            const auto ts3 = tnow();
            seg[0] += secs(ts0, ts1); seg[1] += secs(ts1, ts2); seg[2] += secs(ts2, ts3);
            if ((diag_step + 1) % 50 == 0)
                fprintf(stderr, "[as_diag] segments so far: index+readback %.3f s, factor+sums+GEMM %.3f s, counts %.3f s (columns factored %lld)\n",
                        seg[0], seg[1], seg[2], cols_factored);
            if (!diag_birth.p) {
                KF_CUDA(ctx, diag_birth.ensure((size_t)P * P * sizeof(int) + 8 * sizeof(double)));
                KF_CUDA(ctx, cudaMemsetAsync(diag_birth.p, 0, diag_birth.bytes, st));
            }
            double* d_stats = reinterpret_cast<double*>(diag_birth.as<char>() + (size_t)P * P * sizeof(int));
            kf_as_diag_kernel<<<ncol, 256, 0, st>>>(Am, Bm, SG, GA, GB, C, ld, P, cs, lam, rel, ++diag_step + 1, diag_birth.as<int>(), d_stats);
            if (diag_step % 50 == 0) {
                double hs[6];
                KF_CUDA(ctx, cudaMemcpyAsync(hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
                KF_CUDA(ctx, cudaStreamSynchronize(st));
                static auto t_start = std::chrono::steady_clock::now();
                if (diag_step == 50) t_start = std::chrono::steady_clock::now();
                fprintf(stderr, "[as_diag] +%.3f s ", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
                fprintf(stderr, "[as_diag] step %d: full %.3e flop, with age-ordered reuse %.3e (%.2fx), entering %.0f leaving %.0f, changed columns %.0f of %.0f\n",
                        diag_step, hs[0], hs[1], hs[0] / std::max(hs[1], 1.0), hs[2], hs[3], hs[4], hs[5]);
            }
        }
        kf_as_apply_kernel<<<egrid, 256, 0, st>>>(Am, Bm, SG, GA, GB, C, Kout, ld, P, cs, lam, rel, d_chg);
        KF_CUDA(ctx, cudaGetLastError());
        ctx->launches += 5;
        lam_now = lam;
        lam_prev = lam;
        return KF_OK;
    };
    auto settled = [&]() { return changed == 0 && !clamped; };

    // ---- cold start: the pattern is the largest |C| entry and lam starts at max|C| (K = 0 is optimal there)
    kf_as_absmax_kernel<<<1, 1024, 0, st>>>(C, ld, P, AsCols{0, P, fix_c0, fix_c1}, d_scal);
    double cmax = 0;
    KF_CUDA(ctx, cudaMemcpyAsync(&cmax, d_scal, sizeof(double), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    kf_as_init_kernel<<<egrid, 256, 0, st>>>(C, cmax, SG, ld, P, Pp, cs);
    KF_CUDA(ctx, cudaMemsetAsync(d_chg, 1, sizeof(int) * P, st));        // every column is new (any non-zero value is a flag)
    KF_CUDA(ctx, cudaMemsetAsync(Am, 0, mat * sizeof(double), st));
    KF_CUDA(ctx, cudaMemsetAsync(Bm, 0, mat * sizeof(double), st));
    ctx->launches += 2;
    lam_prev = cmax;
    // ---- the budgets, ascending.  The iteration is a contraction only while the pattern change per step is small enough
    // (config 3a: fine at 5 % of the support, 4 of 64 budgets diverge at 10 %, 13 at 20 %), and the admissible fraction is
    // problem dependent — so it is self-tuning: if the settle phase stops contracting (the number of changes does not
    // decrease over 4 unclamped steps) or the step cap is hit, the budget restarts from the pattern it began with and
    // half the fraction (at most 4 times; the smaller fraction is kept for the following budgets).
    KF_CUDA(ctx, ctx->d_K2.ensure(mat * sizeof(double)));
    double* SG_save = ctx->d_K2.as<double>();
    for (int ob = 0; ob < nb; ++ob) {
        const int b = order[ob];
        double* Kb = K_all + (size_t)b * mat;
        KF_CUDA(ctx, cudaMemcpyAsync(SG_save, SG, mat * sizeof(double), cudaMemcpyDeviceToDevice, st));
        const double lam_save = lam_prev;
        int it = 0, conv = 0;
        for (int attempt = 0; attempt < 5 && !conv; ++attempt) {
            if (attempt) {
                frac = std::max(0.5 * frac, 0.002);
                KF_CUDA(ctx, cudaMemcpyAsync(SG, SG_save, mat * sizeof(double), cudaMemcpyDeviceToDevice, st));
                KF_CUDA(ctx, cudaMemsetAsync(d_chg, 1, sizeof(int) * P, st));      // the pattern was replaced wholesale
                lam_prev = lam_save;
            }
            long long prev_changed = -1;
            int stalled = 0;
            for (int k = 0; k < iter_cap; ++k) {
                KF_TRY(step(t[b], Kb));
                ++it;
                if (settled()) { conv = 1; break; }
                if (!clamped) {
                    stalled = (prev_changed >= 0 && changed >= prev_changed) ? stalled + 1 : 0;
                    prev_changed = changed;
                    if (stalled >= 4) break;          // not contracting: retry with a smaller step
                } else {
                    prev_changed = -1;
                    stalled = 0;
                }
            }
        }
        res[b].iters = it;
        res[b].lam = lam_now;
        res[b].capped = conv ? 0 : 1;
        res[b].l1 = 0;
        res[b].objective = 0;
    }
    diag_birth.release();
    return KF_OK;
}
