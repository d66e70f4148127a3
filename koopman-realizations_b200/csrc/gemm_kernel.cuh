// FP64 tensor-core (DMMA) tile kernel body, templated on the pipeline configuration.
//
//   out[m][n] (+)= alpha * sum_k w[k] * A[m][k] * B[n][k]      (BM x BN tile per CTA)
//
// On sm_100a the FP64 tensor path is the warp-level mma.sync m8n8k4 (SASS DMMA.8x8x4);
// tcgen05 has no f64 kind.  Operands stream global/L2 -> shared through a STAGES-deep
// cp.async (LDGSTS) pipeline, BK contraction elements per stage.
//
// Shared layout: row r of an operand tile holds its BK k-values contiguously, padded.
//   VEC = false: pad to BK+4 doubles; fragment = LDS.64 at (row, kk*4 + lane%4); the 16 lanes of a
//                half-warp hit 16 distinct 8-byte bank pairs ((row*4 + k) mod 16) -> conflict-free.
//   VEC = true : pad to BK+8 doubles; the 4 lanes of a DMMA k-group own k = 2q, 2q+1 of each group of 8,
//                so ONE LDS.128 feeds two consecutive DMMA k-steps (half the shared-load instructions);
//                rows r, r+1 land in disjoint 64-byte halves of the 128-byte bank window -> conflict-free.
#pragma once
#include "kf_internal.h"

namespace kfg {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

template <int BK_, int STAGES_, int WM_, int WN_, bool VEC_, int BM_ = KF_BM, int BN_ = KF_BN, int MINB_ = 1>
struct Cfg {
    static constexpr int BK = BK_, STAGES = STAGES_, WM = WM_, WN = WN_, BM = BM_, BN = BN_, MINB = MINB_;
    static constexpr bool VEC = VEC_;
    static constexpr int THREADS = 32 * WM * WN;
    static constexpr int LDSP = VEC ? BK + 8 : BK + 4;       // padded smem row, doubles
    static constexpr int A_STAGE = BM * LDSP, B_STAGE = BN * LDSP;
    static constexpr int MI = BM / WM / 8, NJ = BN / WN / 8;          // DMMA tiles per warp
    static constexpr int CHUNKS_A = BM * (BK / 2) / THREADS;          // 16-byte chunks per thread, A operand
    static constexpr int CHUNKS_B = BN * (BK / 2) / THREADS;
    static constexpr size_t SMEM = (size_t)STAGES * (A_STAGE + B_STAGE + BK) * sizeof(double);
    static_assert(BM * (BK / 2) % THREADS == 0 && BN * (BK / 2) % THREADS == 0, "loader mapping");
};

template <class C, bool WEIGHTED, bool PREFETCH>
__device__ __forceinline__ void gemm_tile_body(const KfGemmTask& t, double* smem) {
    constexpr int BK = C::BK, STAGES = C::STAGES, LDSP = C::LDSP, MI = C::MI, NJ = C::NJ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / C::WN, wn = warp % C::WN;
    double* As = smem;
    double* Bs = smem + STAGES * C::A_STAGE;
    double* Ws = Bs + STAGES * C::B_STAGE;

    // global->shared assignment: 128 rows x BK/2 sixteen-byte chunks per operand per stage
    constexpr int CPR = BK / 2;   // chunks per row
    const double* a_src[C::CHUNKS_A];
    const double* b_src[C::CHUNKS_B];
    const int l_row = tid / CPR, l_kc = (tid % CPR) * 2;      // chunk i of a thread: row l_row + i * RSTEP
    constexpr int RSTEP = C::THREADS / CPR;
#pragma unroll
    for (int i = 0; i < C::CHUNKS_A; ++i) a_src[i] = t.A + (long long)min(l_row + i * RSTEP, t.a_rows - 1) * t.lda + l_kc;
#pragma unroll
    for (int i = 0; i < C::CHUNKS_B; ++i) b_src[i] = t.B + (long long)min(l_row + i * RSTEP, t.b_rows - 1) * t.ldb + l_kc;
    const int s_off0 = l_row * LDSP + l_kc;
    const int nk = (t.k1 - t.k0) / BK;

    auto load_stage = [&](int stage, int kt) {
        const int k = t.k0 + kt * BK;
        double* as = As + stage * C::A_STAGE;
        double* bs = Bs + stage * C::B_STAGE;
#pragma unroll
        for (int i = 0; i < C::CHUNKS_A; ++i) cp_async16(as + s_off0 + i * RSTEP * LDSP, a_src[i] + k);
#pragma unroll
        for (int i = 0; i < C::CHUNKS_B; ++i) cp_async16(bs + s_off0 + i * RSTEP * LDSP, b_src[i] + k);
        if (WEIGHTED && tid < CPR) cp_async16(Ws + stage * BK + tid * 2, t.W + k + tid * 2);
    };

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    const int fr = (lane >> 2) * LDSP + (C::VEC ? 2 * (lane & 3) : (lane & 3));
    const int a_row0 = wm * (C::BM / C::WM), b_row0 = wn * (C::BN / C::WN);
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nxt = kt + STAGES - 1;
            if (nxt < nk) load_stage(nxt % STAGES, nxt);
            cp_async_commit();
        }
        if (PREFETCH && t.accumulate && t.ldn == 1 && kt == nk - 4) {
            // pull this tile's accumulator lines into L2 ahead of the read-modify-write epilogue
            const char* o = reinterpret_cast<const char*>(t.out);
            constexpr int LPR = C::BN * 8 / 128;   // 128-byte lines per output row
            for (int l = tid; l < C::BM * LPR; l += C::THREADS)
                prefetch_l2(o + ((size_t)(l / LPR) * t.ldm * 8) + (size_t)(l % LPR) * 128);
        }
        const int stage = kt % STAGES;
        const double* as = As + stage * C::A_STAGE + a_row0 * LDSP + fr;
        const double* bs = Bs + stage * C::B_STAGE + b_row0 * LDSP + fr;
        const double* ws = Ws + stage * BK + (C::VEC ? 2 * (lane & 3) : (lane & 3));
        if constexpr (!C::VEC) {
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) {
                double a[MI], b[NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i) a[i] = as[i * 8 * LDSP + kk * 4];
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[j] = bs[j * 8 * LDSP + kk * 4];
                if (WEIGHTED) {
                    const double w = ws[kk * 4];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) b[j] *= w;
                }
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        } else {
#pragma unroll
            for (int k2 = 0; k2 < BK / 8; ++k2) {
                double2 a[MI], b[NJ];
#pragma unroll
                for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const double2*>(as + i * 8 * LDSP + k2 * 8);
#pragma unroll
                for (int j = 0; j < NJ; ++j) b[j] = *reinterpret_cast<const double2*>(bs + j * 8 * LDSP + k2 * 8);
                if (WEIGHTED) {
                    const double2 w = *reinterpret_cast<const double2*>(ws + k2 * 8);
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        b[j].x *= w.x;
                        b[j].y *= w.y;
                    }
                }
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds (m = lane/4, n = 2*(lane%4)+{0,1}) of every 8x8 DMMA tile
    const int m_base = a_row0 + (lane >> 2);
    const int n_base = b_row0 + 2 * (lane & 3);
    const bool vec = (t.ldn == 1);
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int m = m_base + i * 8;
        if (m >= t.a_rows) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int n = n_base + j * 8;
            double v0 = t.alpha * acc[i][j][0], v1 = t.alpha * acc[i][j][1];
            double* p = t.out + (long long)m * t.ldm + (long long)n * t.ldn;
            if (vec && n + 1 < t.b_rows) {
                double2* p2 = reinterpret_cast<double2*>(p);
                if (t.accumulate) {
                    double2 o = *p2;
                    v0 += o.x;
                    v1 += o.y;
                }
                *p2 = make_double2(v0, v1);
            } else {
                if (n < t.b_rows) {
                    if (t.accumulate) v0 += p[0];
                    p[0] = v0;
                }
                if (n + 1 < t.b_rows) {
                    if (t.accumulate) v1 += p[t.ldn];
                    p[t.ldn] = v1;
                }
            }
        }
    }
}

}  // namespace kfg
