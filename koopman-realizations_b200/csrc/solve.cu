// Least-squares solve of the fit on the GPU (replaces `K = Px \ Py`, Ksysid.m:1069).
//
// Gram route (this file, part 1): deterministic split-K slab reduction, assembly of the
// full G = Px'Px and C = Px'Py from the accumulator tiles (expanding the bilinear
// Kronecker blocks), diagonal-pivoted Cholesky of G with a rank tolerance, and the BASIC
// solution on the pivot set (zeros elsewhere) — the semantics of MATLAB's mldivide for a
// rank-deficient tall matrix (QR with column pivoting: the remaining squared column norms
// of QRCP are exactly the Schur-complement diagonal of G, so the pivot order is the same).
// The triangular solves are blocked: 128x128 diagonal blocks in shared memory, everything
// else through the DMMA kernel of gemm.cu.
#include "kf_internal.h"

namespace {

// ---------------------------------------------------------------- split-K slabs
__global__ void kf_reduce_slabs_kernel(double* __restrict__ accum, long long slab, int nsplit) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < slab;
         i += (long long)gridDim.x * blockDim.x) {
        double v = accum[i];
        for (int s = 1; s < nsplit; ++s) {   // fixed order: deterministic
            v += accum[(long long)s * slab + i];
            accum[(long long)s * slab + i] = 0.0;   // folded: a later accumulate / reduce starts from zero
        }
        accum[i] = v;
    }
}

// ---------------------------------------------------------------- assembly
struct AsmArgs {
    int model, N, P, m, Rx, Pp;
    double* G;
    double* C;
};

__global__ void __launch_bounds__(256) kf_assemble_kernel(const double* __restrict__ accum,
                                                          const KfTile* __restrict__ meta, AsmArgs s) {
    const KfTile t = meta[blockIdx.x];
    const double* tile = accum + (long long)blockIdx.x * KF_TILE_ELEMS;
    const long long ld = s.Pp;
    for (int e = threadIdx.x; e < KF_TILE_ELEMS; e += blockDim.x) {
        const int mm = e / KF_BN, nn = e % KF_BN;
        const int r = t.tm * KF_BM + mm, c = t.tn * KF_BN + nn;
        const double v = tile[e];
        if (s.model != KF_BILINEAR) {
            if (t.kind == 0) {
                if (r >= s.Rx || c >= s.Rx || c > r) continue;   // lower triangle only, mirrored: exactly symmetric
                s.G[(long long)c * ld + r] = v;
                s.G[(long long)r * ld + c] = v;
                if (s.model == KF_LINEAR) {   // Py = [psi(y), u]: C(:, N+i) = G(:, N+i)  (Ksysid.m:1063)
                    if (c >= s.N) s.C[(long long)c * ld + r] = v;
                    if (r >= s.N) s.C[(long long)r * ld + c] = v;
                }
            } else {
                if (r >= s.Rx || c >= s.N) continue;
                s.C[(long long)c * ld + r] = v;
            }
        } else {
            if (r >= s.N || c >= s.N) continue;
            const long long ra = (long long)t.a * s.N + r, rb = (long long)t.b * s.N + r;
            const long long ca = (long long)t.a * s.N + c, cb = (long long)t.b * s.N + c;
            if (t.kind == 0) {
                if (c > r) continue;
                // G[(a,r),(b,c)] = sum u_a u_b psi_r psi_c: symmetric in (a<->b) and (r<->c)
                s.G[cb * ld + ra] = v;
                s.G[ca * ld + rb] = v;
                s.G[rb * ld + ca] = v;
                s.G[ra * ld + cb] = v;
            } else {
                // C[(a,r),(b,c)] = sum u_a u_b psix_r psiy_c = C[(b,r),(a,c)]
                s.C[cb * ld + ra] = v;
                s.C[ca * ld + rb] = v;
            }
        }
    }
}

// ---------------------------------------------------------------- pivoted Cholesky
struct PcholState {
    double piv0;     // first (largest) pivot = R11^2
    double minpiv;   // smallest accepted pivot
    int rank;
    int done;
    // parameters (written by the host before the factorisation; the kernels only read them, so one captured graph
    // serves every tolerance and every forced prefix)
    double tol2;     // relative tolerance^2 against `ref`
    double tolabs2;  // absolute tolerance^2 (0: none)
    double ref;      // the pivot the relative tolerance refers to: the first FREE pivot (column `forced`)
    int forced;      // the first `forced` columns are taken in their given order (no search, no swap)
    int pad;
    double rejected; // the pivot candidate the factorisation stopped at (-1: none, every column was accepted)
};

// one CTA: pick the largest remaining diagonal, symmetric swap j<->p, scale column/row j
__global__ void __launch_bounds__(1024) kf_pchol_pivot_kernel(double* W, long long ld, int P, int j, int* perm,
                                                              PcholState* st, double* pivots) {
    if (st->done) return;
    __shared__ double sval[32];
    __shared__ int sidx[32];
    __shared__ int s_p;
    __shared__ double s_d;
    const int tid = threadIdx.x;
    double best = -1.0;
    int bi = P;
    const int jend = (j < st->forced) ? j + 1 : P;   // forced prefix: column j itself
    for (int i = j + tid; i < jend; i += blockDim.x) {
        const double v = W[(long long)i * ld + i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { sval[tid >> 5] = best; sidx[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
        best = (tid < (blockDim.x >> 5)) ? sval[tid] : -1.0;
        bi = (tid < (blockDim.x >> 5)) ? sidx[tid] : P;
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, off);
            const int oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (tid == 0) {
            if (j == 0) st->piv0 = best;
            if (j == st->forced) st->ref = best;
            // a forced column only has to be positive; a free one must clear both tolerances (false for NaN as well)
            const bool ok = (best > 0.0) && (j < st->forced || (best > st->tol2 * st->ref && best > st->tolabs2));
            if (ok && pivots) pivots[j] = best;
            if (!ok) {
                st->rank = j;
                st->done = 1;
                st->rejected = best;
                s_p = -1;
            } else {
                s_p = bi;
                s_d = sqrt(best);
                st->minpiv = best;
                st->rank = j + 1;
            }
        }
    }
    __syncthreads();
    const int p = s_p;
    if (p < 0) return;
    if (p != j) {
        for (int i = tid; i < P; i += blockDim.x) {   // swap columns j <-> p
            const double a = W[(long long)j * ld + i], b = W[(long long)p * ld + i];
            W[(long long)j * ld + i] = b;
            W[(long long)p * ld + i] = a;
        }
        __syncthreads();
        for (int i = tid; i < P; i += blockDim.x) {   // swap rows j <-> p
            const double a = W[(long long)i * ld + j], b = W[(long long)i * ld + p];
            W[(long long)i * ld + j] = b;
            W[(long long)i * ld + p] = a;
        }
        if (tid == 0) {
            const int a = perm[j];
            perm[j] = perm[p];
            perm[p] = a;
        }
        __syncthreads();
    }
    const double d = s_d;
    for (int i = j + tid; i < P; i += blockDim.x) {
        const double l = (i == j) ? d : W[(long long)j * ld + i] / d;
        W[(long long)j * ld + i] = l;   // column j: L(i,j)
        W[(long long)i * ld + j] = l;   // row j:    L(i,j) again (k-contiguous copy for the DMMA operands)
    }
}

// trailing update W(i,k) -= L(i,j) L(k,j), i,k > j (both triangles: stays exactly symmetric)
__global__ void __launch_bounds__(256) kf_pchol_update_kernel(double* W, long long ld, int P, int j,
                                                              const PcholState* st) {
    if (st->done) return;
    const int i = j + 1 + blockIdx.x * 64 + (threadIdx.x & 63);
    const int k0 = j + 1 + blockIdx.y * 16 + (threadIdx.x >> 6) * 4;
    if (i >= P) return;
    const double li = W[(long long)j * ld + i];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = k0 + q;
        if (k < P) {
            const double lk = W[(long long)k * ld + j];
            W[(long long)k * ld + i] = fma(-li, lk, W[(long long)k * ld + i]);
        }
    }
}

// ---- blocked variant (LAPACK dpstrf structure): inside a block of NB columns the factor is built
// left-looking (the Schur-complement diagonal is tracked as W(i,i) - dots[i]), and the trailing matrix is
// updated once per block by the DMMA kernel instead of once per column by a memory-bound rank-1 kernel.
constexpr int PCHOL_NB = 64;

// In-block column step: three small grid-parallel kernels per column (a single-CTA kernel doing the strided
// row swap cost 22 us per column).
// dcur[i] is the current Schur-complement diagonal (contiguous, so the pivot search is coalesced).
struct PcholStep {   // lives right after PcholState in device memory
    int p;           // pivot row of the current column (-1: stop)
    int pad;
    double ljj;      // L(j,j)
};

__global__ void kf_pchol_diag_kernel(const double* __restrict__ W, long long ld, int P, double* dcur) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) dcur[i] = W[(long long)i * ld + i];
}

__global__ void __launch_bounds__(1024) kf_pcholc_argmax_kernel(int P, int j, int* perm, double* dcur, PcholState* st,
                                                                PcholStep* step, double* pivots) {
    if (st->done) {
        if (threadIdx.x == 0) step->p = -1;
        return;
    }
    __shared__ double sval[32];
    __shared__ int sidx[32];
    const int tid = threadIdx.x;
    double best = -1.0;
    int bi = P;
    const int jend = (j < st->forced) ? j + 1 : P;   // forced prefix: column j itself
    for (int i = j + tid; i < jend; i += blockDim.x) {
        const double v = dcur[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { sval[tid >> 5] = best; sidx[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
        best = (tid < (blockDim.x >> 5)) ? sval[tid] : -1.0;
        bi = (tid < (blockDim.x >> 5)) ? sidx[tid] : P;
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, off);
            const int oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (tid == 0) {
            if (j == 0) st->piv0 = best;
            if (j == st->forced) st->ref = best;
            const bool ok = (best > 0.0) && (j < st->forced || (best > st->tol2 * st->ref && best > st->tolabs2));
            if (ok && pivots) pivots[j] = best;
            if (!ok) {
                st->rank = j;
                st->done = 1;
                st->rejected = best;
                step->p = -1;
            } else {
                step->p = bi;
                step->ljj = sqrt(best);
                st->minpiv = best;
                st->rank = j + 1;
                if (bi != j) {
                    const int a = perm[j];
                    perm[j] = perm[bi];
                    perm[bi] = a;
                    const double d = dcur[j];
                    dcur[j] = dcur[bi];
                    dcur[bi] = d;
                }
            }
        }
    }
}

// symmetric permutation j <-> p of the full matrix, one thread per index i
__global__ void __launch_bounds__(256) kf_pcholc_swap_kernel(double* W, long long ld, int P, int j, const PcholStep* step) {
    const int p = step->p;
    if (p < 0 || p == j) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    if (i != j && i != p) {
        double a = W[(long long)j * ld + i], b = W[(long long)p * ld + i];   // column entries (i,j) <-> (i,p)
        W[(long long)j * ld + i] = b;
        W[(long long)p * ld + i] = a;
        a = W[(long long)i * ld + j];                                         // row entries (j,i) <-> (p,i)
        b = W[(long long)i * ld + p];
        W[(long long)i * ld + j] = b;
        W[(long long)i * ld + p] = a;
    } else if (i == j) {                                                      // the 2 x 2 intersection
        double a = W[(long long)j * ld + j], b = W[(long long)p * ld + p];
        W[(long long)j * ld + j] = b;
        W[(long long)p * ld + p] = a;
        a = W[(long long)p * ld + j];
        b = W[(long long)j * ld + p];
        W[(long long)p * ld + j] = b;
        W[(long long)j * ld + p] = a;
    }
}

__global__ void __launch_bounds__(256) kf_pcholc_col_kernel(double* W, long long ld, int P, int j, int j0, double* dcur,
                                                            const PcholStep* step) {
    if (step->p < 0) return;
    const double ljj = step->ljj;
    const int i = j + 1 + blockIdx.x * 256 + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) W[(long long)j * ld + j] = ljj;
    if (i >= P) return;
    double v = W[(long long)j * ld + i];
    for (int k = j0; k < j; ++k) v = fma(-W[(long long)k * ld + i], W[(long long)k * ld + j], v);
    const double l = v / ljj;
    W[(long long)j * ld + i] = l;
    W[(long long)i * ld + j] = l;
    dcur[i] = fma(-l, l, dcur[i]);
}

// zero rows/cols >= rank of the factor so padded contraction ranges contribute nothing
__global__ void kf_pchol_clean_kernel(double* W, long long ld, int Pp, const PcholState* st) {
    const int r = st->rank;
    const long long n = (long long)Pp * Pp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % Pp), k = (int)(e / Pp);
        if (i >= r || k >= r) W[(long long)k * ld + i] = 0.0;
    }
}

// X(i,n) = C(perm[i], n) for i < rank, 0 otherwise
__global__ void kf_gather_rows_kernel(const double* __restrict__ C, long long ldc, const int* __restrict__ perm,
                                      const PcholState* st, int Pp, int ncols, double* X, long long ldx) {
    const int r = st->rank;
    const long long n = (long long)Pp * ncols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % Pp), c = (int)(e / Pp);
        X[(long long)c * ldx + i] = (i < r) ? C[(long long)c * ldc + perm[i]] : 0.0;
    }
}

// K(perm[i], n) = X(i,n) for i < rank; K zeroed beforehand
__global__ void kf_scatter_rows_kernel(const double* __restrict__ X, long long ldx, const int* __restrict__ perm,
                                       const int* __restrict__ rank_ptr, int ncols, double* K, long long ldk) {
    const int r = *rank_ptr;
    const long long n = (long long)r * ncols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % r), c = (int)(e / r);
        K[(long long)c * ldk + perm[i]] = X[(long long)c * ldx + i];
    }
}

// Triangular solve with one 128x128 diagonal block T = W[I:I+128, I:I+128] held in shared memory.
// lower=1: solve L z = y (forward), lower=0: solve L' x = z (backward).  S[i*128+k] = W(I+k, I+i)
// (column I+i of the full symmetric factor): forward uses k >= i (= L(k,i)), backward k <= i (= L(i,k)).
// One warp per right-hand-side column; lane holds rows lane + 32 q.
__global__ void __launch_bounds__(256) kf_trsm_diag_kernel(const double* __restrict__ W, long long ld, int I,
                                                           const int* __restrict__ rank_ptr, double* X, long long ldx,
                                                           int ncols, int lower) {
    extern __shared__ __align__(16) double S[];
    const int r = *rank_ptr;
    const int nb = min(128, r - I);
    if (nb <= 0) return;
    for (int e = threadIdx.x; e < 128 * 128; e += blockDim.x) {
        const int i = e >> 7, k = e & 127;
        S[e] = (i < nb && k < nb) ? W[(long long)(I + i) * ld + I + k] : (i == k ? 1.0 : 0.0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int col = blockIdx.x * 8 + warp; col < ncols; col += gridDim.x * 8) {
        double* x = X + (long long)col * ldx + I;
        double y[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) y[q] = (lane + 32 * q < nb) ? x[lane + 32 * q] : 0.0;
        if (lower) {
#pragma unroll
            for (int qi = 0; qi < 4; ++qi) {
                for (int li = 0; li < 32; ++li) {
                    const int i = qi * 32 + li;
                    if (i >= nb) break;
                    const double zi = __shfl_sync(0xffffffffu, y[qi], li) / S[i * 128 + i];
                    if (lane == li) y[qi] = zi;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int k = lane + 32 * q;
                        if (q >= qi && k > i) y[q] = fma(-S[i * 128 + k], zi, y[q]);
                    }
                }
            }
        } else {
#pragma unroll
            for (int qi = 3; qi >= 0; --qi) {
                for (int li = 31; li >= 0; --li) {
                    const int i = qi * 32 + li;
                    if (i >= nb) continue;
                    const double zi = __shfl_sync(0xffffffffu, y[qi], li) / S[i * 128 + i];
                    if (lane == li) y[qi] = zi;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int k = lane + 32 * q;
                        if (q <= qi && k < i) y[q] = fma(-S[i * 128 + k], zi, y[q]);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (lane + 32 * q < nb) x[lane + 32 * q] = y[q];
    }
}

}  // namespace

int kf_reduce_slabs(kf_ctx* ctx, double* accum, long long slab_elems, int nsplit, cudaStream_t st) {
    if (nsplit <= 1) return KF_OK;
    int grid = (int)std::min<long long>((slab_elems + 255) / 256, (long long)ctx->sm_count * 8);
    kf_reduce_slabs_kernel<<<grid, 256, 0, st>>>(accum, slab_elems, nsplit);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}

int kf_assemble(kf_ctx* ctx, const double* accum, const KfTile* d_meta, int ntiles, const KfLayout& lay, double* G,
                double* C, cudaStream_t st) {
    const size_t bytes = (size_t)lay.Pp * lay.Pp * sizeof(double);
    KF_CUDA(ctx, cudaMemsetAsync(G, 0, bytes, st));
    KF_CUDA(ctx, cudaMemsetAsync(C, 0, bytes, st));
    AsmArgs s{lay.model, lay.N, lay.P, lay.m, lay.Rx, lay.Pp, G, C};
    kf_assemble_kernel<<<ntiles, 256, 0, st>>>(accum, d_meta, s);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}

// blocked forward substitution  L Z = X  on the leading r x r factor in W (mirrored storage), X(i, n) at X[n * ldx + i]
static int trsm_forward(kf_ctx* ctx, const double* W, long long ld, int r, const int* d_rank, double* X, long long ldx, int ncols,
                        cudaStream_t st) {
    KF_CUDA(ctx, kf_ensure_smem(ctx, kf_trsm_diag_kernel, (size_t)128 * 128 * 8));
    const int nblk = (r + 127) / 128;
    const int grid = std::max(1, std::min((ncols + 7) / 8, ctx->sm_count));
    // for block row I: X[I,:] -= L[I,0:I) Z[0:I,:]  then solve the diagonal block
    for (int b = 0; b < nblk; ++b) {
        const int I = b * 128;
        if (I > 0) {
            KfGemmGrid g{};
            g.A = W + (long long)I * ld;   // A[m][k] = L(I+m,k) = W(k, I+m): column I+m, k contiguous
            g.lda = ld;
            g.B = X;                       // B[n][k] = Z(k,n)
            g.ldb = ldx;
            g.out = X + I;                 // out[m][n] = X(I+m, n)
            g.ldm = 1;
            g.ldn = ldx;
            g.m = std::min(128, r - I);
            g.n = ncols;
            g.k0 = 0;
            g.k1 = I;
            g.alpha = -1.0;
            g.accumulate = 1;
            KF_TRY(kf_launch_gemm_grid(ctx, g, st));
        }
        kf_trsm_diag_kernel<<<grid, 256, 128 * 128 * 8, st>>>(W, ld, I, d_rank, X, ldx, ncols, 1);
        KF_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    return KF_OK;
}

// blocked triangular solves  L Z = X, L' X = Z  on the leading rank x rank factor in W
static int trsm_both(kf_ctx* ctx, const double* W, long long ld, int P, int rank_hint, const int* d_rank, double* X,
                     long long ldx, int ncols, cudaStream_t st) {
    // rank_hint: host copy of the rank (bounds the block loops)
    const int r = rank_hint;
    const int nblk = (r + 127) / 128;
    const int grid = std::max(1, std::min((ncols + 7) / 8, ctx->sm_count));
    const int kmax = (int)kf_roundup(r, KF_BK);
    KF_TRY(trsm_forward(ctx, W, ld, r, d_rank, X, ldx, ncols, st));
    // backward: for block row I (last to first): Z[I,:] -= sum_{k>=I+128} L(k, I+m) X(k,:)
    for (int b = nblk - 1; b >= 0; --b) {
        const int I = b * 128;
        if (I + 128 < r) {
            KfGemmGrid g{};
            g.A = W + (long long)I * ld;   // A[m][k] = L(k, I+m) = W(k, I+m) (lower part), k contiguous
            g.lda = ld;
            g.B = X;
            g.ldb = ldx;
            g.out = X + I;
            g.ldm = 1;
            g.ldn = ldx;
            g.m = 128;
            g.n = ncols;
            g.k0 = I + 128;
            g.k1 = kmax;                   // rows >= rank of W and X are zero
            g.alpha = -1.0;
            g.accumulate = 1;
            KF_TRY(kf_launch_gemm_grid(ctx, g, st));
        }
        kf_trsm_diag_kernel<<<grid, 256, 128 * 128 * 8, st>>>(W, ld, I, d_rank, X, ldx, ncols, 0);
        KF_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    return KF_OK;
}

// ---------------------------------------------------------------- refinement basis (multi-level pivoted Cholesky-QR)
namespace {
// Sp(j, i) = S(perm[j], i): row gather of a Pp x Pp column-major matrix (rows >= P stay zero)
__global__ void kf_rf_gather_rows_kernel(const double* __restrict__ S, const int* __restrict__ perm, int P, int Pp, double* Sp) {
    const long long n = (long long)Pp * Pp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(e % Pp), i = (int)(e / Pp);
        Sp[e] = (j < P && i < P) ? S[(long long)i * Pp + perm[j]] : 0.0;
    }
}
// W(k, c) = 0 for k in [r, r16) and all columns c >= r (mirrored storage: the k-contiguous column c) so that a contraction
// over k in [0, r16) sees exactly L(c, 0:r)
__global__ void kf_rf_zero_kpad_kernel(double* W, long long ld, int P, int r, int r16) {
    const int c = r + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P) return;
    for (int k = r; k < r16; ++k) W[(long long)c * ld + k] = 0.0;
}
__global__ void kf_rf_transpose_kernel(const double* __restrict__ A, double* __restrict__ At, int n) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = bx + threadIdx.x, j = by + r;
        tile[r][threadIdx.x] = (i < n && j < n) ? A[(long long)j * n + i] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = by + threadIdx.x, j = bx + r;
        if (i < n && j < n) At[(long long)j * n + i] = tile[threadIdx.x][r];
    }
}
__global__ void kf_rf_identity_kernel(double* S, int P, int Pp) {
    const long long n = (long long)Pp * Pp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(e % Pp), i = (int)(e / Pp);
        S[e] = (i == j && i < P) ? 1.0 : 0.0;
    }
}
// mirror the lower triangle into the upper one (exactly symmetric input for the pivoted Cholesky)
__global__ void kf_rf_symmetrize_kernel(double* G, int Pp) {
    const long long n = (long long)Pp * Pp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % Pp), j = (int)(e / Pp);
        if (i < j) G[e] = G[(long long)i * Pp + j];
    }
}
}  // namespace

int kf_rf_identity(kf_ctx* ctx, double* S, int P, int Pp, cudaStream_t st) {
    kf_rf_identity_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(S, P, Pp);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}
int kf_rf_symmetrize(kf_ctx* ctx, double* G, int Pp, cudaStream_t st) {
    kf_rf_symmetrize_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(G, Pp);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}
int kf_rf_transpose(kf_ctx* ctx, const double* A, double* At, int n, cudaStream_t st) {
    kf_rf_transpose_kernel<<<dim3((n + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, st>>>(A, At, n);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}
int kf_rf_gather_rows(kf_ctx* ctx, const double* S, const int* d_perm, int P, int Pp, double* Sp, cudaStream_t st) {
    kf_rf_gather_rows_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(S, d_perm, P, Pp, Sp);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return KF_OK;
}

// One level of the refinement basis.  W holds the factor of kf_pchol_factor (NOT cleaned) of the Gram of the current
// features Z = Px S', rank r, pivot order perm.  With L_ext = [[L11, 0], [L21, I]] the new basis is
//     S_new = L_ext^-1 (Pi S)          (S, Sp, S_new: Pp x Pp column-major, row = new feature, column = original feature)
// i.e. the first r new features are orthonormalised (Z_new[:r] = L11^-1 Z[perm[:r]]) and the remaining ones are the residuals
// of the not-yet-accepted columns against them (Z_new[r:] = Z[perm[r:]] - L21 Z_new[:r]).  Sp = Pi S is left in `Sp_out`
// and the result is written over `S`.
int kf_rf_update_basis(kf_ctx* ctx, int P, int Pp, double* W, const int* d_perm, int r, double* S, double* Sp_out, cudaStream_t st) {
    const long long ld = Pp;
    PcholState* d_state = ctx->d_misc.as<PcholState>();
    KF_TRY(kf_rf_gather_rows(ctx, S, d_perm, P, Pp, Sp_out, st));
    KF_CUDA(ctx, cudaMemcpyAsync(S, Sp_out, (size_t)Pp * Pp * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (r <= 0) return KF_OK;
    KF_TRY(trsm_forward(ctx, W, ld, r, &d_state->rank, S, ld, P, st));
    if (r < P) {
        const int r16 = (int)kf_roundup(r, KF_BK);
        if (r16 > r) {
            kf_rf_zero_kpad_kernel<<<(P - r + 255) / 256, 256, 0, st>>>(W, ld, P, r, r16);
            KF_CUDA(ctx, cudaGetLastError());
            ctx->launches += 1;
        }
        // S[r:P, :] -= L21 S[0:r, :]     (S rows [r, r16) enter with zero coefficients)
        KfGemmGrid g{};
        g.A = W + (long long)r * ld;   // A[m][k] = W(k, r+m) = L(r+m, k)
        g.lda = ld;
        g.B = S;                       // B[n][k] = S(k, n)
        g.ldb = ld;
        g.out = S + r;                 // out[m][n] = S(r+m, n)
        g.ldm = 1;
        g.ldn = ld;
        g.m = P - r;
        g.n = P;
        g.k0 = 0;
        g.k1 = r16;
        g.alpha = -1.0;
        g.accumulate = 1;
        KF_TRY(kf_launch_gemm_grid(ctx, g, st));
    }
    return KF_OK;
}

// Diagonal-pivoted Cholesky of the symmetric W (Pp x Pp, ld = Pp; overwritten by the mirrored factor, see the kernels).
//   tol     relative tolerance on |R_jj| against the first FREE pivot; tolabs: absolute tolerance on |R_jj| (0: none)
//   forced  the first `forced` columns are pivots in their given order (refinement passes: the columns accepted by the
//           previous level, already orthonormalised)
//   d_perm  pivot order (ints on the device); initialised to the identity here
//   pivots_host (optional, P doubles): the accepted pivots |R_jj|, j < rank;  rejected_out (optional): |R_jj| candidate the
//           factorisation stopped at (-1 if every column was accepted, 0 for a non-positive candidate)
// The rows / columns >= rank of W still hold L21 (rows) and the untouched Schur complement: kf_pchol_solve cleans them.
int kf_pchol_factor(kf_ctx* ctx, int P, int Pp, double* W, double tol, double tolabs, int forced, int* d_perm, int* rank_out,
                    double* min_piv, double* max_piv, double* pivots_host, cudaStream_t st, double* rejected_out) {
    // state block in d_misc: [PcholState | PcholStep]
    KF_CUDA(ctx, ctx->d_misc.ensure(4096));
    PcholState* d_state = ctx->d_misc.as<PcholState>();
    PcholState h0{};
    h0.tol2 = tol * tol;
    h0.tolabs2 = tolabs * tolabs;
    h0.forced = forced;
    h0.rejected = -1.0;
    KF_CUDA(ctx, cudaMemcpyAsync(d_state, &h0, sizeof(h0), cudaMemcpyHostToDevice, st));
    {
        std::vector<int> id(P);
        for (int i = 0; i < P; ++i) id[i] = i;
        KF_CUDA(ctx, cudaMemcpyAsync(d_perm, id.data(), sizeof(int) * P, cudaMemcpyHostToDevice, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));   // id, h0 go out of scope
    }
    const long long ld = Pp;
    KF_CUDA(ctx, ctx->d_K3.ensure((size_t)(4 * Pp + 8) * sizeof(double) + (size_t)(Pp + 4) * sizeof(int)));
    double* dcur = ctx->d_K3.as<double>();
    double* pivots = dcur + 2 * (size_t)Pp;
    if (P <= 2 * PCHOL_NB) {
        // small problems: unblocked right-looking factorisation (2 launches per column)
        for (int j = 0; j < P; ++j) {
            kf_pchol_pivot_kernel<<<1, 1024, 0, st>>>(W, ld, P, j, d_perm, d_state, pivots);
            const int rem = P - j - 1;
            if (rem > 0) {
                dim3 grid((rem + 63) / 64, (rem + 15) / 16);
                kf_pchol_update_kernel<<<grid, 256, 0, st>>>(W, ld, P, j, d_state);
            }
        }
        ctx->launches += 2LL * P;
    } else {
        // blocked (dpstrf structure): left-looking inside a 64-column block, DMMA trailing update per block
        PcholStep* d_step = reinterpret_cast<PcholStep*>(d_state + 1);
        // The loop is launch-bound (3 tiny kernels per column + one DMMA update per block: ~6,000 launches at P = 4096, no
        // host decision anywhere — pivots, ranks AND the tolerances / forced prefix live in device state), so it is captured
        // ONCE into a CUDA graph per (buffers, P) and replayed by every later factorisation.
        auto enqueue = [&]() -> int {
            for (int j0 = 0; j0 < P; j0 += PCHOL_NB) {
                const int j1 = std::min(P, j0 + PCHOL_NB);
                kf_pchol_diag_kernel<<<(P + 255) / 256, 256, 0, st>>>(W, ld, P, dcur);   // Schur diagonal at block start
                for (int j = j0; j < j1; ++j) {
                    kf_pcholc_argmax_kernel<<<1, 1024, 0, st>>>(P, j, d_perm, dcur, d_state, d_step, pivots);
                    kf_pcholc_swap_kernel<<<(P + 255) / 256, 256, 0, st>>>(W, ld, P, j, d_step);
                    const int rem = P - j - 1;
                    kf_pcholc_col_kernel<<<std::max(1, (rem + 255) / 256), 256, 0, st>>>(W, ld, P, j, j0, dcur, d_step);
                }
                ctx->launches += 1 + 3LL * (j1 - j0);
                if (j1 < P) {
                    KfGemmGrid g{};
                    g.A = W + (long long)j1 * ld;   // A[m][c] = W(c, j1+m) = L(j1+m, c): mirrored copy, c contiguous
                    g.lda = ld;
                    g.B = W + (long long)j1 * ld;
                    g.ldb = ld;
                    g.out = W + (long long)j1 * ld + j1;   // out[m][n] = W(j1+m, j1+n)
                    g.ldm = 1;
                    g.ldn = ld;
                    g.m = P - j1;
                    g.n = P - j1;
                    g.k0 = j0;
                    g.k1 = j1;               // j1 - j0 = 64 for every block that has a trailing part
                    g.alpha = -1.0;
                    g.accumulate = 1;
                    KF_TRY(kf_launch_gemm_grid(ctx, g, st));
                }
            }
            return KF_OK;
        };
        KfPcholGraph& G = ctx->pchol_graph;
        const bool same = G.exec && G.W == W && G.perm == d_perm && G.dcur == dcur && G.state == d_state && G.P == P && G.Pp == Pp &&
                          G.stream == st;
        if (!ctx->opt_graphs) {
            KF_TRY(enqueue());
        } else {
            if (!same) {
                if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
                const long long l0 = ctx->launches;
                const double f0 = ctx->dmma_flops;
                cudaGraph_t graph = nullptr;
                KF_CUDA(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                const int rc = enqueue();
                const cudaError_t ce = cudaStreamEndCapture(st, &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                KF_CUDA(ctx, ce);
                KF_CUDA(ctx, cudaGraphInstantiate(&G.exec, graph, 0));
                cudaGraphDestroy(graph);
                G.W = W; G.perm = d_perm; G.dcur = dcur; G.state = d_state; G.P = P; G.Pp = Pp; G.stream = st;
                G.launches = ctx->launches - l0;
                G.flops = ctx->dmma_flops - f0;
                ctx->launches = l0;                      // counted per replay below
                ctx->dmma_flops = f0;
            }
            KF_CUDA(ctx, cudaGraphLaunch(G.exec, st));
            ctx->launches += G.launches;
            ctx->dmma_flops += G.flops;
        }
    }
    KF_CUDA(ctx, cudaGetLastError());
    PcholState hs;
    KF_CUDA(ctx, cudaMemcpyAsync(&hs, d_state, sizeof(hs), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    const int r = hs.rank;
    if (pivots_host && r > 0) {
        KF_CUDA(ctx, cudaMemcpyAsync(pivots_host, pivots, sizeof(double) * r, cudaMemcpyDeviceToHost, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
        for (int j = 0; j < r; ++j) pivots_host[j] = sqrt(pivots_host[j] > 0 ? pivots_host[j] : 0.0);
    }
    if (r < forced) {
        ctx->err = "pivoted Cholesky broke down inside the forced prefix (column " + std::to_string(r) + " of " +
                   std::to_string(forced) + "): the refinement basis lost positive definiteness";
        return KF_ENUMERIC;
    }
    *rank_out = r;
    *max_piv = sqrt(hs.piv0 > 0 ? hs.piv0 : 0.0);
    *min_piv = sqrt(hs.minpiv > 0 ? hs.minpiv : 0.0);
    if (rejected_out) *rejected_out = hs.rejected > 0 ? sqrt(hs.rejected) : (hs.rejected < 0 ? -1.0 : 0.0);
    return KF_OK;
}

// Basic solution from the factor of kf_pchol_factor: K(perm[0:r], :) = (L11 L11')^-1 C(perm[0:r], :), zeros elsewhere.
// scatter = 0 leaves the solution in PIVOT order (rows 0..r-1 of K) instead of scattering it by perm.
int kf_pchol_solve(kf_ctx* ctx, int P, int Pp, int ncols, double* W, const double* C, double* K, const int* d_perm, int r, int scatter,
                   cudaStream_t st) {
    const long long ld = Pp;
    PcholState* d_state = ctx->d_misc.as<PcholState>();
    kf_pchol_clean_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(W, ld, Pp, d_state);
    // X = C(perm[0:r], :) in the K buffer's scratch twin (d_tmp), solve, scatter
    KF_CUDA(ctx, ctx->d_tmp.ensure((size_t)Pp * Pp * sizeof(double)));
    double* X = scatter ? ctx->d_tmp.as<double>() : K;
    kf_gather_rows_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(C, ld, d_perm, d_state, Pp, ncols, X, ld);
    KF_CUDA(ctx, cudaGetLastError());
    if (scatter) KF_CUDA(ctx, cudaMemsetAsync(K, 0, (size_t)Pp * Pp * sizeof(double), st));
    if (r > 0) {
        KF_TRY(trsm_both(ctx, W, ld, P, r, &d_state->rank, X, ld, ncols, st));
        if (scatter) {
            kf_scatter_rows_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(X, ld, d_perm, &d_state->rank, ncols, K, ld);
            KF_CUDA(ctx, cudaGetLastError());
        }
    }
    ctx->launches += 4;
    return KF_OK;
}

int kf_solve_gram_ls(kf_ctx* ctx, int P, int Pp, int ncols, double* W, const double* C, double* K, double tol, int* d_perm,
                     int* rank_out, double* min_piv, double* max_piv, cudaStream_t st) {
    KF_TRY(kf_pchol_factor(ctx, P, Pp, W, tol, 0.0, 0, d_perm, rank_out, min_piv, max_piv, nullptr, st, nullptr));
    return kf_pchol_solve(ctx, P, Pp, ncols, W, C, K, d_perm, *rank_out, 1, st);
}

// =====================================================================================
// Part 2 — Householder QR with column pivoting on the materialised regressors: the same
// algorithm LAPACK dgeqp3 / MATLAB mldivide run for `Px \ Py` (Ksysid.m:1069, 1216),
// used when [Px | Py] fits in device memory.  Unblocked: per column one single-CTA kernel
// (pivot search on exactly recomputed remaining column norms, column swap, dlarfg
// reflector) and one grid-wide kernel that applies the reflector to every trailing
// column of [A | B] and recomputes the remaining norms.  Rank is decided like MATLAB:
// |R_jj| <= max(size(A)) * eps(|R_11|)  ->  basic solution on the first `rank` pivots.
// =====================================================================================
namespace {

struct QrState {
    double r11;     // |R_11|
    double tol;     // max(M,P) * eps(|R_11|)
    double minpiv;  // last accepted |R_jj|
    double tau;     // current reflector
    int rank;
    int done;
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
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

__global__ void __launch_bounds__(256) kf_qr_colnorms_kernel(const double* __restrict__ A, long long ld, long long M,
                                                             double* vn) {
    __shared__ double sh[32];
    const double* col = A + (long long)blockIdx.x * ld;
    double s = 0.0;
    for (long long i = threadIdx.x; i < M; i += blockDim.x) s = fma(col[i], col[i], s);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) vn[blockIdx.x] = s;
}

// one CTA: pivot, swap, reflector (dlarfg) for column j
__global__ void __launch_bounds__(1024) kf_qr_pivot_kernel(double* A, long long ld, long long M, long long Mtol, int P, int j, double* vn,
                                                           int* perm, QrState* st) {
    if (st->done) return;
    __shared__ double sh[32];
    __shared__ double sval[32];
    __shared__ int sidx[32];
    __shared__ int s_p;
    const int tid = threadIdx.x;
    double best = -1.0;
    int bi = P;
    for (int i = j + tid; i < P; i += blockDim.x) {
        const double v = vn[i];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, off);
        const int oi = __shfl_down_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { sval[tid >> 5] = best; sidx[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
        best = (tid < (blockDim.x >> 5)) ? sval[tid] : -1.0;
        bi = (tid < (blockDim.x >> 5)) ? sidx[tid] : P;
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, best, off);
            const int oi = __shfl_down_sync(0xffffffffu, bi, off);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (tid == 0) {
            const double nrm = sqrt(best > 0.0 ? best : 0.0);
            if (j == 0) {
                st->r11 = nrm;
                // eps(x) = distance to the next double above x
                const double e = (nrm > 0.0) ? (__longlong_as_double(__double_as_longlong(nrm) + 1) - nrm) : 0.0;
                st->tol = (double)(Mtol > P ? Mtol : P) * e;      // max(size(A)) * eps(|R_11|), the rows of the ORIGINAL matrix
            }
            if (!(nrm > st->tol)) {
                st->rank = j;
                st->done = 1;
                s_p = -1;
            } else {
                s_p = bi;
                st->rank = j + 1;
                st->minpiv = nrm;
            }
        }
    }
    __syncthreads();
    const int p = s_p;
    if (p < 0) return;
    double* cj = A + (long long)j * ld;
    if (p != j) {
        double* cp = A + (long long)p * ld;
        for (long long i = tid; i < M; i += blockDim.x) {
            const double a = cj[i];
            cj[i] = cp[i];
            cp[i] = a;
        }
        if (tid == 0) {
            const int a = perm[j];
            perm[j] = perm[p];
            perm[p] = a;
            vn[p] = vn[j];
        }
        __syncthreads();
    }
    // dlarfg on x = A[j:, j]
    double s = 0.0;
    for (long long i = j + 1 + tid; i < M; i += blockDim.x) s = fma(cj[i], cj[i], s);
    const double xnorm2 = block_sum(s, sh);
    const double alpha = cj[j];
    double tau = 0.0, beta = alpha, scale = 0.0;
    if (xnorm2 > 0.0) {
        beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
    }
    __syncthreads();
    for (long long i = j + 1 + tid; i < M; i += blockDim.x) cj[i] *= scale;
    if (tid == 0) {
        cj[j] = beta;
        st->tau = tau;
    }
}

// apply H_j = I - tau v v' (v = [1; A[j+1:, j]]) to column c = j+1+blockIdx.x of [A | B];
// recompute the remaining norm of the column (rows > j) for columns of A.
__global__ void __launch_bounds__(256) kf_qr_apply_kernel(double* A, long long ld, long long M, int P, int j, double* vn,
                                                          const QrState* st) {
    if (st->done) return;
    __shared__ double sh[32];
    const int c = j + 1 + blockIdx.x;
    const double* v = A + (long long)j * ld;
    double* col = A + (long long)c * ld;
    const double tau = st->tau;
    double s = 0.0;
    for (long long i = j + 1 + threadIdx.x; i < M; i += blockDim.x) s = fma(v[i], col[i], s);
    const double w = tau * (block_sum(s, sh) + col[j]);
    double nn = 0.0;
    for (long long i = j + 1 + threadIdx.x; i < M; i += blockDim.x) {
        const double x = fma(-w, v[i], col[i]);
        col[i] = x;
        nn = fma(x, x, nn);
    }
    __syncthreads();
    if (threadIdx.x == 0) col[j] -= w;
    if (c < P) {
        nn = block_sum(nn, sh);
        if (threadIdx.x == 0) vn[c] = nn;
    }
}

// W (Pp x Pp, zeroed) <- R11 mirrored: W(i,k) = W(k,i) = R(i,k), i <= k < rank
__global__ void kf_qr_extract_r_kernel(const double* __restrict__ A, long long ld, const QrState* st, double* W, long long ldw) {
    const int r = st->rank;
    const long long n = (long long)r * r;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % r), k = (int)(e / r);
        if (i <= k) {
            const double v = A[(long long)k * ld + i];
            W[(long long)k * ldw + i] = v;
            W[(long long)i * ldw + k] = v;
        }
    }
}

// X(i, c) = (Q'B)(i, c) = AB(i, P + c) for i < rank, else 0
__global__ void kf_qr_gather_rhs_kernel(const double* __restrict__ AB, long long ld, int P, int Pc, const QrState* st, double* X,
                                        long long ldx, int Pp) {
    const int r = st->rank;
    const long long n = (long long)Pp * Pc;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % Pp), c = (int)(e / Pp);
        X[(long long)c * ldx + i] = (i < r) ? AB[(long long)(P + c) * ld + i] : 0.0;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Stage 1 of the blocked QRCP route: Householder QR WITHOUT pivoting of the tall matrix, 128 columns at a time, the trailing
// matrix and the right-hand sides updated with the compact WY form on the DMMA GEMMs:  A = Q1 R1.  Column pivoting then only
// has to look at R1 (P x P): R1 Pi = Q2 R2 gives A Pi = (Q1 Q2) R2, the same factorisation — column norms and every later
// pivot decision are invariant under the orthogonal Q1 — with the BLAS-2 part on a matrix that no longer depends on M.
constexpr int BQ_NB = 128;         // largest panel width

// reflector of column j (rows j .. M-1) by one CTA: dlarfg; tau[j] = 0 for a zero tail
__device__ void bqr_house(double* cj, long long M, int j, double* tau, double* sh) {
    const int tid = threadIdx.x;
    double s = 0.0;
    for (long long i = j + 1 + tid; i < M; i += blockDim.x) s = fma(cj[i], cj[i], s);
    const double xnorm2 = block_sum(s, sh);
    const double alpha = cj[j];
    double t = 0.0, beta = alpha, scale = 0.0;
    if (xnorm2 > 0.0) {
        beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
        t = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
    }
    __syncthreads();
    for (long long i = j + 1 + tid; i < M; i += blockDim.x) cj[i] *= scale;
    if (tid == 0) { cj[j] = beta; tau[j] = t; }
}
__global__ void __launch_bounds__(1024) kf_bqr_house_kernel(double* A, long long ld, long long M, int j, double* tau) {
    __shared__ double sh[32];
    bqr_house(A + (long long)j * ld, M, j, tau, sh);
}
// apply H_j to the panel columns j+1 .. jend-1 (one CTA each); the CTA of column j+1 then forms ITS reflector, so a panel
// costs one launch per column
__global__ void __launch_bounds__(1024) kf_bqr_step_kernel(double* A, long long ld, long long M, int j, double* tau, int do_house) {
    __shared__ double sh[32];
    const int c = j + 1 + blockIdx.x;
    const double* v = A + (long long)j * ld;
    double* col = A + (long long)c * ld;
    const double t = tau[j];
    if (t != 0.0) {
        double s = 0.0;
        for (long long i = j + 1 + threadIdx.x; i < M; i += blockDim.x) s = fma(v[i], col[i], s);
        const double w = t * (block_sum(s, sh) + col[j]);
        for (long long i = j + 1 + threadIdx.x; i < M; i += blockDim.x) col[i] = fma(-w, v[i], col[i]);
        __syncthreads();
        if (threadIdx.x == 0) col[j] -= w;
        __syncthreads();
    }
    if (blockIdx.x == 0 && do_house) bqr_house(col, M, c, tau, sh);
}
// explicit V of a panel: Vp (rows x nb, ld = rows) for the matrix rows q0 .. (q0 = p0 rounded down to a multiple of 64, roff = p0 - q0):
// unit lower trapezoidal below row roff, zero above; columns >= jb zero
__global__ void kf_bqr_vpanel_kernel(const double* __restrict__ A, long long ld, int p0, int roff, int jb, int nb, long long rows, double* Vp) {
    const long long n = rows * nb;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e % rows - roff;             // row relative to p0
        const int i = (int)(e / rows);
        double v = 0.0;
        if (i < jb && r >= i) v = (r == i) ? 1.0 : A[(long long)(p0 + i) * ld + p0 + r];
        Vp[e] = v;
    }
}
// T factor of the panel (dlarft, forward / columnwise): T(0:j, j) = -tau_j T(0:j, 0:j) S(0:j, j), S = Vp' Vp; T upper triangular
__global__ void __launch_bounds__(BQ_NB) kf_bqr_tfactor_kernel(const double* __restrict__ S, const double* __restrict__ tau, int jb, int nb, double* T) {
    extern __shared__ double sT[];      // [nb][nb + 1]
    const int i = threadIdx.x;
    const int lds = nb + 1;
    for (int k = 0; k < nb; ++k) sT[i * lds + k] = 0.0;
    __syncthreads();
    for (int j = 0; j < jb; ++j) {
        const double tj = tau[j];
        double acc = 0.0;
        if (i < j) {
            for (int k = i; k < j; ++k) acc = fma(sT[i * lds + k], S[(long long)j * nb + k], acc);
            acc *= -tj;
        } else if (i == j) acc = tj;
        __syncthreads();
        if (i <= j) sT[i * lds + j] = acc;
        __syncthreads();
    }
    for (int k = 0; k < nb; ++k) T[(long long)k * nb + i] = sT[i * lds + k];   // column-major T
}
// zero the part of the first `P` columns below the diagonal (the Householder vectors) within the top `rows` rows
__global__ void kf_bqr_clear_lower_kernel(double* A, long long ld, int P, int rows) {
    const long long n = (long long)P * rows;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e % rows), c = (int)(e / rows);
        if (r > c) A[(long long)c * ld + r] = 0.0;
    }
}

}  // namespace

// backward substitution only:  U X = Z with U = L' stored mirrored in W
static int trsm_backward(kf_ctx* ctx, const double* W, long long ld, int r, const int* d_rank, double* X, long long ldx, int ncols,
                         cudaStream_t st) {
    KF_CUDA(ctx, kf_ensure_smem(ctx, kf_trsm_diag_kernel, (size_t)128 * 128 * 8));
    const int nblk = (r + 127) / 128;
    const int grid = std::max(1, std::min((ncols + 7) / 8, ctx->sm_count));
    const int kmax = (int)kf_roundup(r, KF_BK);
    for (int b = nblk - 1; b >= 0; --b) {
        const int I = b * 128;
        if (I + 128 < r) {
            KfGemmGrid g{};
            g.A = W + (long long)I * ld;
            g.lda = ld;
            g.B = X;
            g.ldb = ldx;
            g.out = X + I;
            g.ldm = 1;
            g.ldn = ldx;
            g.m = 128;
            g.n = ncols;
            g.k0 = I + 128;
            g.k1 = kmax;
            g.alpha = -1.0;
            g.accumulate = 1;
            KF_TRY(kf_launch_gemm_grid(ctx, g, st));
        }
        kf_trsm_diag_kernel<<<grid, 256, 128 * 128 * 8, st>>>(W, ld, I, d_rank, X, ldx, ncols, 0);
        KF_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    return KF_OK;
}

// stage 1 (see above): on return the top min(M, P) rows of AB hold [R1 | Q1' B]; *rows_out = rows the pivoted stage works on
static int bqr_stage1(kf_ctx* ctx, long long Mp, int P, int Pc, double* AB, long long ld, int* rows_out, cudaStream_t st) {
    const int ncols = P + Pc;
    const int nfac = (int)std::min<long long>(Mp, P);
    const int nb = (ctx->opt_qr_nb == 32 || ctx->opt_qr_nb == 64 || ctx->opt_qr_nb == 128) ? ctx->opt_qr_nb : 64;
    // V'A2 and V'V contract over all rows: cut into chunks of ~256 rows summed in a second level (a single sequential
    // accumulation over 1e4 rows cost a digit on config 2a / 2b: 7e-10 instead of 1e-10 against the extended-precision solution)
    const int kmax = std::max(1, std::min(ctx->opt_qr_ksplit, 64));
    KF_CUDA(ctx, ctx->d_bqr.ensure(((size_t)Mp * nb + (size_t)nb * nb * (kmax + 1) + (size_t)nb * ncols * (kmax + 1) + (size_t)P + 64) * sizeof(double)));
    double* Vp = ctx->d_bqr.as<double>();
    double* S = Vp + (size_t)Mp * nb;                    // kmax slabs
    double* T = S + (size_t)nb * nb * kmax;
    double* W = T + (size_t)nb * nb;                     // kmax slabs
    double* W2 = W + (size_t)nb * ncols * kmax;
    double* tau = W2 + (size_t)nb * ncols;
    const size_t tsm = (size_t)nb * (nb + 1) * sizeof(double);
    KF_CUDA(ctx, kf_ensure_smem(ctx, kf_bqr_tfactor_kernel, (size_t)BQ_NB * (BQ_NB + 1) * sizeof(double)));
    for (int p0 = 0; p0 < nfac; p0 += nb) {
        const int jb = std::min(nb, nfac - p0);
        if (!(ctx->opt_qr_blocked == 2 && p0 > 0)) kf_bqr_house_kernel<<<1, 1024, 0, st>>>(AB, ld, Mp, p0, tau);
        if (ctx->opt_qr_blocked == 2) {                  // diagnostic: every reflector applied to ALL remaining columns, no compact WY
            for (int j = p0; j < p0 + jb; ++j)
                if (ncols - j - 1 > 0) kf_bqr_step_kernel<<<ncols - j - 1, 1024, 0, st>>>(AB, ld, Mp, j, tau, (j + 1 < nfac) ? 1 : 0);
            ctx->launches += jb;
            continue;
        }
        for (int j = p0; j + 1 < p0 + jb; ++j) kf_bqr_step_kernel<<<p0 + jb - j - 1, 1024, 0, st>>>(AB, ld, Mp, j, tau, 1);
        ctx->launches += jb;
        const int c0 = p0 + jb, nc = ncols - c0;
        if (nc <= 0) break;
        const int q0 = p0 / 64 * 64, roff = p0 - q0;     // operand rows start at a multiple of 64 (GEMM tile / alignment); Vp is zero above p0
        const long long rows = Mp - q0;
        kf_bqr_vpanel_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(AB, ld, p0, roff, jb, nb, rows, Vp);
        KfGemmGrid g{};
        g.A = Vp; g.lda = rows; g.B = Vp; g.ldb = rows; g.out = S; g.ldm = 1; g.ldn = nb;         // S = Vp' Vp
        g.m = nb; g.n = nb; g.k0 = 0; g.k1 = (int)rows; g.alpha = 1.0; g.accumulate = 0;
        const int nz = kf_gemm_ksplit((int)rows, std::min<long long>(kmax, std::max<long long>(1, rows / 256)));
        g.ksplit = nz; g.slab = (long long)nb * nb;
        KF_TRY(kf_launch_gemm_grid(ctx, g, st));
        KF_TRY(kf_reduce_slabs(ctx, S, g.slab, nz, st));
        kf_bqr_tfactor_kernel<<<1, nb, tsm, st>>>(S, tau + p0, jb, nb, T);
        g.A = Vp; g.lda = rows; g.B = AB + (long long)c0 * ld + q0; g.ldb = ld; g.out = W; g.ldm = 1; g.ldn = nb;   // W = Vp' A2
        g.m = nb; g.n = nc; g.k0 = 0; g.k1 = (int)rows;
        g.ksplit = nz; g.slab = (long long)nb * nc;
        KF_TRY(kf_launch_gemm_grid(ctx, g, st));
        KF_TRY(kf_reduce_slabs(ctx, W, g.slab, nz, st));
        g.ksplit = 0; g.slab = 0;
        g.A = T; g.lda = nb; g.B = W; g.ldb = nb; g.out = W2; g.ldm = 1; g.ldn = nb;              // W2 = T' W
        g.m = nb; g.n = nc; g.k0 = 0; g.k1 = nb;
        KF_TRY(kf_launch_gemm_grid(ctx, g, st));
        g.A = W2; g.lda = nb; g.B = Vp; g.ldb = rows; g.out = AB + (long long)c0 * ld + q0; g.ldm = ld; g.ldn = 1;   // A2 -= Vp W2
        g.m = nc; g.n = (int)rows; g.k0 = 0; g.k1 = nb; g.alpha = -1.0; g.accumulate = 1;
        KF_TRY(kf_launch_gemm_bkmajor(ctx, g, st));
        ctx->launches += 2;
    }
    KF_CUDA(ctx, cudaGetLastError());
    kf_bqr_clear_lower_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(AB, ld, P, nfac);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    *rows_out = nfac;
    return KF_OK;
}

int kf_solve_qr_ls(kf_ctx* ctx, long long M, int P, int Pc, double* AB, long long ldab, double* X, long long ldx, int* d_perm,
                   int* rank_out, double* min_piv, double* max_piv, cudaStream_t st) {
    // X: Pp x Pc output (column-major, ld = ldx >= Pp), rows = original column index of A
    const int Pp = (int)kf_roundup(P, KF_BM);
    KF_CUDA(ctx, ctx->d_misc.ensure(1024 + (size_t)Pp * sizeof(int)));
    KF_CUDA(ctx, ctx->d_K3.ensure((size_t)Pp * sizeof(double)));
    QrState* d_state = reinterpret_cast<QrState*>(ctx->d_misc.as<char>() + 512);
    double* vn = ctx->d_K3.as<double>();
    QrState h0{};
    KF_CUDA(ctx, cudaMemcpyAsync(d_state, &h0, sizeof(h0), cudaMemcpyHostToDevice, st));
    {
        std::vector<int> id(P);
        for (int i = 0; i < P; ++i) id[i] = i;
        KF_CUDA(ctx, cudaMemcpyAsync(d_perm, id.data(), sizeof(int) * P, cudaMemcpyHostToDevice, st));
        KF_CUDA(ctx, cudaStreamSynchronize(st));
    }
    // tall problems: blocked unpivoted QR first, then the pivoted factorisation only sees R1 (P rows)
    long long Mw = M;                                    // rows the pivoted stage works on
    const long long Mp = kf_qr_ld(M);
    if (ctx->opt_qr_blocked && ldab == Mp && P >= 2 * BQ_NB && M >= (long long)P + P / 2) {
        int rows = 0;
        KF_TRY(bqr_stage1(ctx, Mp, P, Pc, AB, ldab, &rows, st));
        Mw = rows;
    }
    kf_qr_colnorms_kernel<<<P, 256, 0, st>>>(AB, ldab, Mw, vn);
    const int steps = (int)std::min<long long>(Mw, P);
    const int ncols = P + Pc;
    for (int j = 0; j < steps; ++j) {
        kf_qr_pivot_kernel<<<1, 1024, 0, st>>>(AB, ldab, Mw, M, P, j, vn, d_perm, d_state);
        if (ncols - j - 1 > 0) kf_qr_apply_kernel<<<ncols - j - 1, 256, 0, st>>>(AB, ldab, Mw, P, j, vn, d_state);
    }
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2LL * steps + 1;
    QrState hs;
    KF_CUDA(ctx, cudaMemcpyAsync(&hs, d_state, sizeof(hs), cudaMemcpyDeviceToHost, st));
    KF_CUDA(ctx, cudaStreamSynchronize(st));
    const int r = hs.rank;
    *rank_out = r;
    *max_piv = hs.r11;
    *min_piv = hs.minpiv;
    // R11 -> mirrored Pp x Pp factor in d_W; rhs -> d_tmp; back-substitute; scatter to X
    const size_t mat = (size_t)Pp * Pp * sizeof(double);
    const int Pcp = (int)kf_roundup(Pc, KF_BN);
    KF_CUDA(ctx, ctx->d_W.ensure(mat));
    KF_CUDA(ctx, ctx->d_tmp.ensure((size_t)Pp * Pcp * sizeof(double)));
    double* W = ctx->d_W.as<double>();
    double* Z = ctx->d_tmp.as<double>();
    KF_CUDA(ctx, cudaMemsetAsync(W, 0, mat, st));
    KF_CUDA(ctx, cudaMemsetAsync(Z, 0, (size_t)Pp * Pcp * sizeof(double), st));
    KF_CUDA(ctx, cudaMemset2DAsync(X, (size_t)ldx * sizeof(double), 0, (size_t)Pp * sizeof(double), Pc, st));
    if (r > 0) {
        kf_qr_extract_r_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(AB, ldab, d_state, W, Pp);
        kf_qr_gather_rhs_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(AB, ldab, P, Pc, d_state, Z, Pp, Pp);
        KF_CUDA(ctx, cudaGetLastError());
        KF_TRY(trsm_backward(ctx, W, Pp, r, &d_state->rank, Z, Pp, Pc, st));
        kf_scatter_rows_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(Z, Pp, d_perm, &d_state->rank, Pc, X, ldx);
        KF_CUDA(ctx, cudaGetLastError());
        ctx->launches += 3;
    }
    return KF_OK;
}
