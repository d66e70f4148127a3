// FP64-exact Gram / cross-covariance on the INT8 tensor cores ("Ozaki scheme II": integer modular arithmetic + CRT).
//
// The DMMA kernel (gemm.cu) already fills the FP64 tensor pipe to ~96 % — the only way past that pipe is the 5th-generation
// tensor cores, which have no f64 kind.  Scheme: per chunk of Mc snapshots
//   1. row exponents e_r = ceil(log2 max_k |z_r[k]|) of the lifted panel rows                        (oz_rowmax_kernel)
//   2. fixed point  a_r[k] = rint(z_r[k] * 2^(s - e_r)),  |a| <= 2^s, s = 51;  residues a mod p_t as int8 for T = 15 pairwise
//      coprime moduli p_t <= 256 (centred, [-128, 127]); the bilinear rows u_a psi_j are formed on the fly     (oz_residue_kernel)
//   3. T INT8 GEMMs with INT32 accumulation in TMEM:  c_t = (A_t B_t') mod p_t   (oz_gemm.cuh: tcgen05.mma kind::i8, TMA)
//   4. CRT:  x = sum_k a_m[k] b_n[k]  EXACTLY, because |x| <= Mc 2^(2s) < prod(p_t) / 2:  x / M = frac( sum_t c_t y_t / p_t ),
//      evaluated as two EXACT 41-bit fixed-point sums in FP64, then  G[m][n] += x 2^(e_m + e_n - 2s)              (oz_crt_kernel)
// Every product z_m[k] z_n[k] of the rounded operands is summed exactly; the only roundings are the 51-bit fixed-point
// conversion of the operands (relative to the row maximum) and one FP64 rounding of the chunk's contribution — measured
// 8e-17 relative Frobenius error against an extended-precision Gram, i.e. BELOW plain FP64 accumulation (3.8e-16).
// 15 INT8 GEMMs replace one FP64 GEMM; at the measured 2.96 POP/s that is 3.7x the DMMA kernel's rate on the same panel.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "kf_internal.h"
#include "oz_gemm.cuh"
#include "tma_host.h"

namespace {

constexpr int OZ_T = 15;
constexpr int OZ_S = 51;
// pairwise coprime, <= 256.  255 is left out on purpose: for every p <= 253 the centred residue |r| <= p/2 + 1 fits an int8
// as it is, and for p = 256 the wrap of the byte IS the reduction, so the residue kernel needs no range fix.
const unsigned OZ_MODS[OZ_T] = {256, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193};
__constant__ int OZ_MODS_D[OZ_T] = {256, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193};

struct OzConst {
    double inv_p[OZ_T];          // 1 / p_t
    double pd[OZ_T];             // p_t
    // CRT weights y_t / p_t (y_t = (M / p_t)^-1 mod p_t) as two 41-bit fixed-point pieces held in doubles:
    // y_t / p_t = v1 + v2 (+ < 2^-82), v1 on the 2^-41 grid, v2 on the 2^-82 grid.  A residue (8 bits) times a piece and the
    // sum of 15 such products stay within 53 bits, so both sums are EXACT in FP64; the neglected tail is an absolute error of
    // M 2^-70 ~ 2^47 in x, 14 bits BELOW the rounding floor 2^-53 sum|a||b| ~ 2^61 of an FP64 dot product of the same operands.
    double v1[OZ_T], v2[OZ_T];
    double m_hi, m_lo;           // prod p_t as a double-double
};
constexpr double OZ_MAGIC = 6755399441055744.0;   // 1.5 * 2^52: (x + MAGIC) - MAGIC = rint(x) for |x| < 2^51

// ------------------------------------------------------------------ 1. row exponents
// rows: list of panel rows; e[i] = exponent with max |row| <= 2^e
__global__ void __launch_bounds__(256) oz_rowmax_kernel(const double* __restrict__ panel, long long ld, int Mc, const int* __restrict__ rows,
                                                        int* __restrict__ e) {
    __shared__ double red[8];
    const double* r = panel + (long long)rows[blockIdx.x] * ld;
    double m = 0.0;
    for (int k = threadIdx.x; k < Mc; k += 256) m = fmax(m, fabs(r[k]));
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmax(m, red[w]);
        int ex = 0;
        if (m > 0.0 && isfinite(m)) frexp(m, &ex);      // m = f 2^ex, f in [0.5, 1)  ->  m <= 2^ex
        e[blockIdx.x] = ex;
    }
}

// ------------------------------------------------------------------ 2. residues
struct OzResArgs {
    const double* psi;       // first panel row of this side's features (row j at psi + j * ld)
    const double* uw;        // weight rows: u_a at uw + wrow[a] * ld (bilinear); unused otherwise
    long long ld;
    int Mc;
    int nfeat;               // feature rows of this side (NX: N for bilinear, all X rows otherwise)
    int nblk;                // Kronecker blocks: m + 1 for bilinear, 1 otherwise
    int rows;                // regressor rows of this side that are needed (<= nfeat * nblk): rows beyond are not written
    int wrow[8];             // panel row (relative to uw) of u_a, a = 1 .. nblk-1  (index 0 unused)
    const int* e_feat;       // exponents of the feature rows
    const int* e_u;          // exponents of u_a (index 0 = 0)
    int8_t* out;             // [T][rows][Mc]
    OzConst c;
};

__global__ void __launch_bounds__(128) oz_residue_kernel(const OzResArgs a) {
    const int kg = blockIdx.x * 128 + threadIdx.x;       // group of 16 snapshots
    const int j = blockIdx.y;
    const int k0 = kg * 16;
    if (k0 >= a.Mc) return;
    double z[16];
    const double2* src = reinterpret_cast<const double2*>(a.psi + (long long)j * a.ld + k0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const double2 v = src[q];
        z[2 * q] = v.x;
        z[2 * q + 1] = v.y;
    }
    const int ej = a.e_feat[j];
    for (int blk = 0; blk < a.nblk; ++blk) {
        const int row = blk * a.nfeat + j;
        if (row >= a.rows) break;
        double f[16];
        int ex = ej;
        if (blk == 0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) f[q] = z[q];
        } else {
            const double2* us = reinterpret_cast<const double2*>(a.uw + (long long)a.wrow[blk] * a.ld + k0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double2 u = us[q];
                f[2 * q] = __dmul_rn(u.x, z[2 * q]);          // the regressor entry u_a psi_j, rounded as the reference forms it
                f[2 * q + 1] = __dmul_rn(u.y, z[2 * q + 1]);
            }
            ex += a.e_u[blk];
        }
        const double sc = ldexp(1.0, OZ_S - ex);
        int xl[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            f[q] = rint(f[q] * sc);                               // exact integer, |f| <= 2^51
            xl[q] = __double2loint(f[q] + OZ_MAGIC);              // its low 32 bits (f + MAGIC is exact: both integers, sum < 2^53)
        }
        const size_t plane = (size_t)a.rows * a.Mc;
        int8_t* dst = a.out + (size_t)row * a.Mc + k0;
#pragma unroll 1
        for (int t = 0; t < OZ_T; ++t) {
            const double ip = a.c.inv_p[t];
            const int pi = OZ_MODS_D[t];
            uint32_t w[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t pk = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    // q = rint(x / p) up to +-1, read as an integer from the low mantissa bits of x / p + MAGIC; the residue
                    // r = x - q p is small (|r| <= p/2 + 1), so its low 32 bits follow from the low 32 bits of x and q:
                    // ONE FP64 instruction and one integer multiply-add per (element, modulus)
                    const int ql = __double2loint(fma(f[g * 4 + e], ip, OZ_MAGIC));
                    const int r = xl[g * 4 + e] - ql * pi;
                    pk |= ((uint32_t)r & 0xFFu) << (8 * e);         // p <= 253: r in [-127, 127]; p = 256: the byte wrap is the reduction
                }
                w[g] = pk;
            }
            *reinterpret_cast<uint4*>(dst + (size_t)t * plane) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// ------------------------------------------------------------------ 4. CRT + accumulate
struct OzCrtArgs {
    const uint8_t* res;              // [T][plane]: G plane rows x LD, then C plane rows x LD (row-major, n contiguous)
    unsigned long long plane;        // bytes per modulus
    unsigned long long c_off;        // offset of the C plane inside a modulus plane
    int LD;
    int xrows, yrows;                // regressor rows of X / of Y (the C columns)
    int nfx, nfy;                    // feature rows per Kronecker block (exponent lookup)
    const int* e_fx; const int* e_fy; const int* e_u;
    double* accG; double* accC;      // row-major [m * Pp + n] accumulators (G: n <= m only)
    int Pp;
    int symC;                        // the C symmetry is in use as well
    int symN;                        // > 0: Kronecker block size N of a bilinear regressor whose block symmetry is exploited:
                                     //   G block (a, b): only local rows >= local columns are computed (the block is symmetric),
                                     //   C block (a, b): only a <= b (C_(a,b) = C_(b,a));  0: everything on / below the diagonal
    OzConst c;
};

// grid: (column groups of 1024, rows, 2 planes); thread = 4 consecutive n of one row m.
// x / M = frac(sum_t c_t y_t / p_t): the two fixed-point sums are exact, the fraction is centred (|x| < M / 4 by construction),
// and x = M f carries ~2 ulp relative to |x| ITSELF (small entries — nearly orthogonal rows — keep their relative accuracy).
__global__ void __launch_bounds__(256) oz_crt_kernel(const OzCrtArgs a) {
    const int m = blockIdx.y;
    const int n0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    const bool isC = blockIdx.z != 0;
    const int ncols = isC ? a.yrows : a.xrows;
    if (n0 >= ncols || (!isC && n0 > m)) return;      // the Gram is computed on / below the diagonal only
    if (a.symN) {                                     // whole 4-column group outside the computed part of its Kronecker block
        const int ba = m / a.symN, bb = n0 / a.symN, r = m % a.symN, c0 = n0 % a.symN;     // (4 | N: a group never straddles blocks)
        if (isC ? (a.symC && ba > bb) : (c0 > r)) return;
    }
    const uint8_t* src = a.res + (isC ? a.c_off : 0ull) + (size_t)m * a.LD + n0;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < OZ_T; ++t) {
        const uint32_t pk = *reinterpret_cast<const uint32_t*>(src + (size_t)t * a.plane);
        const double v1 = a.c.v1[t], v2 = a.c.v2[t];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double c = __hiloint2double(0x43300000, (int)((pk >> (8 * e)) & 0xFFu)) - 4503599627370496.0;   // byte -> double, exact
            s1[e] = fma(c, v1, s1[e]);
            s2[e] = fma(c, v2, s2[e]);
        }
    }
    const int em = a.e_fx[m % a.nfx] + a.e_u[m / a.nfx];
    double* acc = (isC ? a.accC : a.accG) + (size_t)m * a.Pp + n0;
    double v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int n = n0 + e;
        double x = 0.0;
        if (n < ncols && (isC || n <= m) && (!a.symN || isC || (n % a.symN) <= (m % a.symN))) {   // computed entries only
            const int en = isC ? (a.e_fy[n % a.nfy] + a.e_u[n / a.nfy]) : (a.e_fx[n % a.nfx] + a.e_u[n / a.nfx]);
            const double f1 = s1[e] - ((s1[e] + OZ_MAGIC) - OZ_MAGIC);       // centred fraction of the leading sum (exact)
            const double f = f1 + s2[e];
            x = ldexp(fma(f, a.c.m_hi, f * a.c.m_lo), em + en - 2 * OZ_S);
        }
        v[e] = x;
    }
    double2* p2 = reinterpret_cast<double2*>(acc);
    double2 o0 = p2[0], o1 = p2[1];
    o0.x += v[0]; o0.y += v[1]; o1.x += v[2]; o1.y += v[3];
    p2[0] = o0;
    p2[1] = o1;
}

// fill the entries the block symmetry skipped (row-major accumulators): G block (a >= b): local upper = transposed local lower;
// C block (a > b) = C block (b, a)
__global__ void oz_mirror_kernel(double* __restrict__ accG, double* __restrict__ accC, int Pp, int P, int N, int doC) {
    const long long n_el = (long long)P * P;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(e / P), n = (int)(e % P);
        const int ba = m / N, bb = n / N, r = m % N, c = n % N;
        if (ba >= bb && c > r) accG[(size_t)m * Pp + n] = accG[(size_t)(ba * N + c) * Pp + bb * N + r];
        if (doC && ba > bb) accC[(size_t)m * Pp + n] = accC[(size_t)(bb * N + r) * Pp + ba * N + c];
    }
}

__global__ void oz_add_kernel(double* __restrict__ a, const double* __restrict__ b, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] += b[i];
}
// linear model: Py = [psi(y), u]  ->  C(:, N + i) = G(:, N + i)  (Ksysid.m:1063); G, C column-major Pp x Pp
__global__ void oz_linear_cols_kernel(const double* __restrict__ G, double* __restrict__ C, int Pp, int P, int N, int m) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P) return;
    for (int i = 0; i < m; ++i) C[(size_t)(N + i) * Pp + r] = G[(size_t)(N + i) * Pp + r];
}

int make_map3(CUtensorMap* out, void* base, unsigned long long K, unsigned long long rows, unsigned long long T, unsigned box_rows) {
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
        fn = reinterpret_cast<encode_fn>(p);
    }
    const cuuint64_t gdim[3] = {K, rows, T};
    const cuuint64_t gstr[2] = {K, K * rows};
    const cuuint32_t box[3] = {128, box_rows, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}

OzConst make_consts() {
    OzConst c{};
    unsigned __int128 M = 1;
    for (int t = 0; t < OZ_T; ++t) M *= OZ_MODS[t];
    for (int t = 0; t < OZ_T; ++t) {
        const unsigned p = OZ_MODS[t];
        c.inv_p[t] = 1.0 / (double)p;
        c.pd[t] = (double)p;
        unsigned mp = 1;                              // (M / p_t) mod p_t
        for (int s = 0; s < OZ_T; ++s)
            if (s != t) mp = (unsigned)(((unsigned long long)mp * (OZ_MODS[s] % p)) % p);
        unsigned y = 1;
        while ((unsigned long long)mp * y % p != 1) ++y;      // inverse by search (p <= 256)
        // y / p to 82 fraction bits, cut into two 41-bit pieces
        const unsigned __int128 Nf = ((unsigned __int128)y << 82) / p;
        const unsigned long long mask = (1ull << 41) - 1;
        c.v1[t] = std::ldexp((double)(unsigned long long)(Nf >> 41), -41);
        c.v2[t] = std::ldexp((double)((unsigned long long)Nf & mask), -82);
    }
    c.m_hi = (double)M;
    const unsigned __int128 mh = (unsigned __int128)c.m_hi;
    c.m_lo = M >= mh ? (double)(M - mh) : -(double)(mh - M);
    return c;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
struct KfOzState {
    bool ready = false;
    int xrows = 0, yrows = 0, LD = 0, Mc = 0, nfx = 0, nfy = 0, nblk = 1, Pp = 0;
    int symN = 0;             // Kronecker block symmetry in use (bilinear, N a multiple of 256)
    bool symC = false;        // ... for the cross product too (all of Py wanted)
    unsigned long long plane = 0, c_off = 0;
    KfBuf d_rx[KF_MAX_PIPES], d_ry[KF_MAX_PIPES], d_res[KF_MAX_PIPES], d_exp[KF_MAX_PIPES], d_rowlist, d_tasks;
    CUtensorMap mxa[KF_MAX_PIPES], mxb[KF_MAX_PIPES], myb[KF_MAX_PIPES];
    oz::Params prm{};
    int ntasks = 0;
    int nrowlist = 0;
    OzConst c;
};

static void oz_release(KfOzState* s) {
    for (int b = 0; b < KF_MAX_PIPES; ++b) { s->d_rx[b].release(); s->d_ry[b].release(); s->d_res[b].release(); s->d_exp[b].release(); }
    s->d_rowlist.release();
    s->d_tasks.release();
}
void kf_oz_destroy(kf_ctx* ctx) {
    if (ctx->oz) {
        oz_release(ctx->oz);
        delete ctx->oz;
        ctx->oz = nullptr;
    }
}

bool kf_oz_supported(const KfLayout& L) {
    const int m1 = L.model == KF_BILINEAR ? L.m + 1 : 1;
    return L.nw == 0 && m1 <= 8 && L.Mc % 128 == 0 && L.Mc <= 8192;
}

// buffers, tensor maps and the task list for the layout (called from make_layout when the INT8 engine is selected)
int kf_oz_prepare(kf_ctx* ctx, KfLayout& L) {
    if (!ctx->oz) ctx->oz = new KfOzState();
    KfOzState& S = *ctx->oz;
    S.c = make_consts();
    const bool bil = L.model == KF_BILINEAR;
    S.nblk = bil ? L.m + 1 : 1;
    S.nfx = bil ? L.N : L.Rx;
    S.nfy = L.N;
    S.xrows = L.P;                                         // linear: N + m = Rx; nonlinear: N; bilinear: N (m+1)
    S.yrows = bil ? std::min(L.Pc, L.P) : std::min(L.Pc, L.N);
    S.Mc = L.Mc;
    S.Pp = L.Pp;
    S.LD = (int)kf_roundup(std::max(S.xrows, S.yrows), 256);
    const unsigned long long xr128 = (unsigned long long)kf_roundup(S.xrows, 128);
    S.c_off = xr128 * S.LD;
    S.plane = 2ull * xr128 * S.LD;
    // panel rows whose exponents are needed: X features, Y features, u_a (bilinear)
    std::vector<int> rows;
    for (int j = 0; j < S.nfx; ++j) rows.push_back(L.x_off + j);
    for (int j = 0; j < S.nfy; ++j) rows.push_back(L.y_off + j);
    for (int a = 1; a < S.nblk; ++a) rows.push_back(L.w_off + a);     // pair_index(0, a, m) = a: the weight row u_a
    S.nrowlist = (int)rows.size();
    KF_CUDA(ctx, S.d_rowlist.ensure(rows.size() * sizeof(int)));
    KF_CUDA(ctx, cudaMemcpyAsync(S.d_rowlist.p, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    // Kronecker structure of the bilinear regressor Px = [1; u] (x) psi: G block (a, b) = sum u_a u_b psi psi' is itself symmetric and
    // C block (a, b) = C block (b, a) — 15.3 N^2 instead of 24.5 N^2 tile area (the same saving the DMMA path takes with its
    // weighted blocks).  Needs whole tiles per block (N a multiple of 256); C only when all of Py is wanted.
    S.symN = (bil && L.N % 256 == 0 && ctx->opt_oz_sym) ? L.N : 0;
    const bool symC = S.symN && S.yrows == S.xrows;
    std::vector<oz::Task> tasks;
    for (int t = 0; t < OZ_T; ++t) {
        for (int mt = 0; mt * 128 < S.xrows; ++mt)
            for (int nt = 0; nt * 256 <= mt * 128 + 127 && nt * 256 < S.xrows; ++nt) {
                if (S.symN && (nt * 256) % S.symN > (mt * 128) % S.symN + 127) continue;        // local upper tile of a block
                tasks.push_back({mt * 128, nt * 256, t, 0, (unsigned long long)mt * 128 * S.LD + (unsigned long long)nt * 256});
            }
        for (int mt = 0; mt * 128 < S.xrows; ++mt)
            for (int nt = 0; nt * 256 < S.yrows; ++nt) {
                if (symC && (mt * 128) / S.symN > (nt * 256) / S.symN) continue;               // block a > b: mirrored from (b, a)
                tasks.push_back({mt * 128, nt * 256, t, 1, S.c_off + (unsigned long long)mt * 128 * S.LD + (unsigned long long)nt * 256});
            }
    }
    S.symC = symC;
    S.ntasks = (int)tasks.size();
    KF_CUDA(ctx, S.d_tasks.ensure(tasks.size() * sizeof(oz::Task)));
    KF_CUDA(ctx, cudaMemcpyAsync(S.d_tasks.p, tasks.data(), tasks.size() * sizeof(oz::Task), cudaMemcpyHostToDevice, ctx->stream));
    for (int b = 0; b < L.npipes; ++b) {
        KF_CUDA(ctx, S.d_rx[b].ensure((size_t)OZ_T * S.xrows * S.Mc));
        KF_CUDA(ctx, S.d_ry[b].ensure((size_t)OZ_T * S.yrows * S.Mc));
        KF_CUDA(ctx, S.d_res[b].ensure((size_t)OZ_T * S.plane));
        KF_CUDA(ctx, S.d_exp[b].ensure((size_t)(S.nrowlist + 8) * sizeof(int)));
        KF_CUDA(ctx, cudaMemsetAsync(S.d_exp[b].p, 0, (size_t)(S.nrowlist + 8) * sizeof(int), ctx->stream));
        if (make_map3(&S.mxa[b], S.d_rx[b].p, S.Mc, S.xrows, OZ_T, 128) || make_map3(&S.mxb[b], S.d_rx[b].p, S.Mc, S.xrows, OZ_T, 256) ||
            make_map3(&S.myb[b], S.d_ry[b].p, S.Mc, S.yrows, OZ_T, 256)) {
            ctx->err = "cuTensorMapEncodeTiled failed for the INT8 residue planes";
            return KF_ECUDA;
        }
    }
    KF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // rows / tasks host vectors go out of scope
    oz::Params& p = S.prm;
    p = oz::Params{};
    p.tasks = S.d_tasks.as<oz::Task>();
    p.ntasks = S.ntasks;
    p.K = S.Mc;
    p.plane = S.plane;
    p.ld_out = S.LD;
    p.m_valid = (int)xr128;
    p.n_valid_x = S.LD;
    p.n_valid_y = S.LD;
    for (int t = 0; t < OZ_T; ++t) {
        p.p[t] = OZ_MODS[t];
        p.magic[t] = ((1ull << 37) + OZ_MODS[t] - 1) / OZ_MODS[t];
        p.offset[t] = (int)(((1u << 27) + OZ_MODS[t] - 1) / OZ_MODS[t] * OZ_MODS[t]);
    }
    KF_CUDA(ctx, kf_ensure_smem(ctx, oz::oz_gemm_kernel, oz::SMEM_BYTES));
    // the element-wise kernels of one chunk pipeline should run BESIDE the other pipeline's contraction (which holds 197 KB of
    // shared memory per SM): ask for the same shared-memory carveout so that no SM has to be reconfigured between them
    cudaFuncSetAttribute(oz_rowmax_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(oz_residue_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(oz_crt_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
    S.ready = true;
    return KF_OK;
}

// One chunk: the lifted panel of pipeline b (already written by the lift on stream st) -> exponents -> residues -> INT8 GEMMs
// -> CRT accumulate into this pipeline's dense accumulator set (accG, accC row-major, Pp x Pp each).
// KF_OZ_TRACE=1: CUDA-event timeline of the first chunks (which kernels of the two pipelines overlap) printed by kf_oz_finish
static std::vector<cudaEvent_t> g_tr_ev;
static std::vector<int> g_tr_tag;
static int g_tr_on = -1;
static void tr_mark(int tag, cudaStream_t st) {
    if (g_tr_on < 0) g_tr_on = getenv("KF_OZ_TRACE") ? 1 : 0;
    if (!g_tr_on || g_tr_ev.size() > 400) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    g_tr_ev.push_back(e);
    g_tr_tag.push_back(tag);
}
static void tr_dump() {
    if (g_tr_on != 1 || g_tr_ev.empty()) return;
    cudaDeviceSynchronize();
    for (size_t i = 0; i < g_tr_ev.size(); ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, g_tr_ev[0], g_tr_ev[i]);
        printf("TRACE pipe %d stage %d t=%.3f ms\n", g_tr_tag[i] / 10, g_tr_tag[i] % 10, ms);
    }
    for (cudaEvent_t e : g_tr_ev) cudaEventDestroy(e);
    g_tr_ev.clear();
    g_tr_tag.clear();
    g_tr_on = 0;
}

int kf_oz_chunk(kf_ctx* ctx, const KfLayout& L, int b, const double* panel, double* accG, double* accC, cudaStream_t st, cudaEvent_t ev0,
                cudaEvent_t ev1) {
    KfOzState& S = *ctx->oz;
    tr_mark(b * 10 + 0, st);
    int* e_all = S.d_exp[b].as<int>();
    // exponent table layout: [e_u (8 ints: index 0 stays 0)] [X features] [Y features]
    int* e_u = e_all;
    int* e_fx = e_all + 8;
    int* e_fy = e_fx + S.nfx;
    // the row list is X features, Y features, u_1..u_m: write u exponents to e_u[1..]
    oz_rowmax_kernel<<<S.nfx + S.nfy, 256, 0, st>>>(panel, L.Mc, L.Mc, S.d_rowlist.as<int>(), e_fx);
    if (S.nblk > 1) oz_rowmax_kernel<<<S.nblk - 1, 256, 0, st>>>(panel, L.Mc, L.Mc, S.d_rowlist.as<int>() + S.nfx + S.nfy, e_u + 1);
    OzResArgs ra{};
    ra.ld = L.Mc;
    ra.Mc = L.Mc;
    ra.nblk = S.nblk;
    ra.uw = panel + (long long)L.w_off * L.Mc;
    for (int a = 1; a < S.nblk; ++a) ra.wrow[a] = a;
    ra.e_u = e_u;
    ra.c = S.c;
    const dim3 gx((L.Mc / 16 + 127) / 128, S.nfx), gy((L.Mc / 16 + 127) / 128, S.nfy);
    tr_mark(b * 10 + 1, st);
    ra.psi = panel + (long long)L.x_off * L.Mc; ra.nfeat = S.nfx; ra.rows = S.xrows; ra.e_feat = e_fx; ra.out = S.d_rx[b].as<int8_t>();
    oz_residue_kernel<<<gx, 128, 0, st>>>(ra);
    ra.psi = panel + (long long)L.y_off * L.Mc; ra.nfeat = S.nfy; ra.rows = S.yrows; ra.e_feat = e_fy; ra.out = S.d_ry[b].as<int8_t>();
    oz_residue_kernel<<<gy, 128, 0, st>>>(ra);
    oz::Params p = S.prm;
    p.out = S.d_res[b].as<uint8_t>();
    tr_mark(b * 10 + 2, st);
    if (ev0) KF_CUDA(ctx, cudaEventRecord(ev0, st));       // sampled launches: CUDA events tightly around the INT8 contraction
    oz::oz_gemm_kernel<<<std::min(ctx->sm_count, S.ntasks), oz::THREADS, oz::SMEM_BYTES, st>>>(S.mxa[b], S.mxb[b], S.myb[b], p);
    if (ev1) KF_CUDA(ctx, cudaEventRecord(ev1, st));
    tr_mark(b * 10 + 3, st);
    OzCrtArgs ca{};
    ca.res = S.d_res[b].as<uint8_t>();
    ca.plane = S.plane;
    ca.c_off = S.c_off;
    ca.LD = S.LD;
    ca.xrows = S.xrows; ca.yrows = S.yrows; ca.nfx = S.nfx; ca.nfy = S.nfy;
    ca.e_fx = e_fx; ca.e_fy = e_fy; ca.e_u = e_u;
    ca.accG = accG; ca.accC = accC; ca.Pp = S.Pp;
    ca.symN = S.symN;
    ca.symC = S.symC ? 1 : 0;
    ca.c = S.c;
    oz_crt_kernel<<<dim3((S.LD / 4 + 255) / 256, S.xrows, 2), 256, 0, st>>>(ca);
    tr_mark(b * 10 + 4, st);
    KF_CUDA(ctx, cudaGetLastError());
    ctx->launches += 5 + (S.nblk > 1 ? 1 : 0);
    ctx->i8_ops += 2.0 * oz::BM * oz::BN * (double)L.Mc * S.ntasks;
    ctx->i8_ops_per_launch = 2.0 * oz::BM * oz::BN * (double)L.Mc * S.ntasks;
    return KF_OK;
}

// acc[0] += acc[1] (the two chunk pipelines), then the row-major accumulators become the column-major G (symmetric) and C
int kf_oz_finish(kf_ctx* ctx, const KfLayout& L, double* acc0, double* acc1, cudaStream_t st) {
    tr_dump();
    const long long n = 2LL * L.Pp * L.Pp;
    oz_add_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(acc0, acc1, n);
    KF_CUDA(ctx, cudaGetLastError());
    KF_CUDA(ctx, cudaMemsetAsync(acc1, 0, (size_t)n * sizeof(double), st));      // folded: a second call adds zeros (idempotent)
    ctx->launches += 1;
    return KF_OK;
}
int kf_oz_to_gc(kf_ctx* ctx, const KfLayout& L, double* acc, double* G, double* C, cudaStream_t st) {
    KfOzState& S = *ctx->oz;
    if (S.symN) {
        // the mirror reads entries it never writes (computed ones) and writes entries it never reads: safe in place
        oz_mirror_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(acc, acc + (size_t)L.Pp * L.Pp, L.Pp, L.P, S.symN, S.symC ? 1 : 0);
        KF_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    KF_TRY(kf_rf_transpose(ctx, acc, G, L.Pp, st));                         // row-major lower -> column-major lower
    KF_TRY(kf_rf_symmetrize(ctx, G, L.Pp, st));
    KF_TRY(kf_rf_transpose(ctx, acc + (size_t)L.Pp * L.Pp, C, L.Pp, st));
    if (L.model == KF_LINEAR) {
        oz_linear_cols_kernel<<<(L.P + 255) / 256, 256, 0, st>>>(G, C, L.Pp, L.P, L.N, L.m);
        KF_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    return KF_OK;
}
