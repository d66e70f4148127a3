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

// one pipeline stage of DMMAs: acc += sum over the BK k-values of the stage of (A row) x (w * B row)
template <class C, bool WEIGHTED>
__device__ __forceinline__ void mma_stage(const double* as, const double* bs, const double* ws,
                                          double (&acc)[C::MI][C::NJ][2]) {
    constexpr int BK = C::BK, LDSP = C::LDSP, MI = C::MI, NJ = C::NJ;
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

// epilogue: thread holds (m = lane/4, n = 2*(lane%4)+{0,1}) of every 8x8 DMMA tile
template <class C>
__device__ __forceinline__ void store_tile(const KfGemmTask& t, const double (&acc)[C::MI][C::NJ][2], int a_row0, int b_row0,
                                           int lane) {
    const int m_base = a_row0 + (lane >> 2);
    const int n_base = b_row0 + 2 * (lane & 3);
    const bool vec = (t.ldn == 1);
#pragma unroll
    for (int i = 0; i < C::MI; ++i) {
        const int m = m_base + i * 8;
        if (m >= t.a_rows) continue;
#pragma unroll
        for (int j = 0; j < C::NJ; ++j) {
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

template <class C>
__device__ __forceinline__ void prefetch_out_tile(const KfGemmTask& t, int tid) {
    // pull this tile's accumulator lines into L2 ahead of the read-modify-write epilogue
    const char* o = reinterpret_cast<const char*>(t.out);
    constexpr int LPR = C::BN * 8 / 128;   // 128-byte lines per output row
    for (int l = tid; l < C::BM * LPR; l += C::THREADS)
        prefetch_l2(o + ((size_t)(l / LPR) * t.ldm * 8) + (size_t)(l % LPR) * 128);
}

// ------------------------------------------------------------------------------------------------
// Body 1: per-thread cp.async (LDGSTS) pipeline, one __syncthreads per stage.
template <class C, bool WEIGHTED, bool PREFETCH>
__device__ __forceinline__ void gemm_tile_body(const KfGemmTask& t, double* smem) {
    constexpr int BK = C::BK, STAGES = C::STAGES, LDSP = C::LDSP, MI = C::MI, NJ = C::NJ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / C::WN, wn = warp % C::WN;
    double* As = smem;
    double* Bs = smem + STAGES * C::A_STAGE;
    double* Ws = Bs + STAGES * C::B_STAGE;

    // global->shared assignment: rows x BK/2 sixteen-byte chunks per operand per stage
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
        if (PREFETCH && t.accumulate && t.ldn == 1 && kt == nk - 4) prefetch_out_tile<C>(t, tid);
        const int stage = kt % STAGES;
        mma_stage<C, WEIGHTED>(As + stage * C::A_STAGE + a_row0 * LDSP + fr, Bs + stage * C::B_STAGE + b_row0 * LDSP + fr,
                               Ws + stage * BK + (C::VEC ? 2 * (lane & 3) : (lane & 3)), acc);
    }
    cp_async_wait<0>();
    store_tile<C>(t, acc, a_row0, b_row0, lane);
}

// ------------------------------------------------------------------------------------------------
// Body 1b: same pipeline, but the B operand is K-MAJOR:  out[m][n] (+)= alpha * sum_k A[m][k] * B[k][n]  with B stored
// B[k * ldb + n] (n contiguous) — a plain row-major "NN" product.  Used by the refinement pass of the Gram route
// (solve.cu / api.cu): Z = S * Px, where the lifted panel Px is feature-major (row = feature k, column = snapshot n),
// i.e. the contraction index is NOT contiguous in it.  B tile in shared memory: [BK][BN + 2] doubles; the lane (g, q) of a
// DMMA needs B[k = 8 k2 + 2q (+1)][n = g]: rows 2q apart are 2 * 66 * 8 B = 32 B (mod 128) apart and g spans 32 B, so a
// half-warp's LDS.64 hit 16 distinct 8-byte banks -> conflict-free.  Requires b_rows == BN (n padded by the caller).
template <class C>
__device__ __forceinline__ void gemm_tile_body_bkmajor(const KfGemmTask& t, double* smem) {
    static_assert(C::VEC, "A fragments use the LDS.128 layout");
    constexpr int BK = C::BK, STAGES = C::STAGES, LDSP = C::LDSP, MI = C::MI, NJ = C::NJ, BN = C::BN;
    constexpr int LDN = BN + 2;
    constexpr int B_STAGE = BK * LDN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / C::WN, wn = warp % C::WN;
    double* As = smem;
    double* Bs = smem + STAGES * C::A_STAGE;

    constexpr int CPR = BK / 2;
    const double* a_src[C::CHUNKS_A];
    const int l_row = tid / CPR, l_kc = (tid % CPR) * 2;
    constexpr int RSTEP = C::THREADS / CPR;
#pragma unroll
    for (int i = 0; i < C::CHUNKS_A; ++i) a_src[i] = t.A + (long long)min(l_row + i * RSTEP, t.a_rows - 1) * t.lda + l_kc;
    const int s_off0 = l_row * LDSP + l_kc;
    // B: BK rows x BN/2 sixteen-byte chunks per stage
    constexpr int BCPR = BN / 2;
    constexpr int CHUNKS_B = BK * BCPR / C::THREADS;
    static_assert(BK * BCPR % C::THREADS == 0, "B loader mapping");
    const int nk = (t.k1 - t.k0) / BK;

    auto load_stage = [&](int stage, int kt) {
        const int k = t.k0 + kt * BK;
        double* as = As + stage * C::A_STAGE;
        double* bs = Bs + stage * B_STAGE;
#pragma unroll
        for (int i = 0; i < C::CHUNKS_A; ++i) cp_async16(as + s_off0 + i * RSTEP * LDSP, a_src[i] + k);
#pragma unroll
        for (int i = 0; i < CHUNKS_B; ++i) {
            const int c = tid + i * C::THREADS, kr = c / BCPR, nc = (c % BCPR) * 2;
            cp_async16(bs + kr * LDN + nc, t.B + (long long)(k + kr) * t.ldb + nc);
        }
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
    const int g = lane >> 2, q = lane & 3;
    const int a_row0 = wm * (C::BM / C::WM), b_row0 = wn * (BN / C::WN);
    const int fra = g * LDSP + 2 * q;
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nxt = kt + STAGES - 1;
            if (nxt < nk) load_stage(nxt % STAGES, nxt);
            cp_async_commit();
        }
        const int stage = kt % STAGES;
        const double* as = As + stage * C::A_STAGE + a_row0 * LDSP + fra;
        const double* bs = Bs + stage * B_STAGE + b_row0 + g;
#pragma unroll
        for (int k2 = 0; k2 < BK / 8; ++k2) {
            double2 a[MI];
            double b0[NJ], b1[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const double2*>(as + i * 8 * LDSP + k2 * 8);
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                b0[j] = bs[(k2 * 8 + 2 * q) * LDN + j * 8];
                b1[j] = bs[(k2 * 8 + 2 * q + 1) * LDN + j * 8];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b0[j]);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b1[j]);
        }
    }
    cp_async_wait<0>();
    store_tile<C>(t, acc, a_row0, b_row0, lane);
}

template <class C>
constexpr size_t bkmajor_smem_bytes() {
    return (size_t)C::STAGES * (C::A_STAGE + C::BK * (C::BN + 2)) * sizeof(double);
}

// ------------------------------------------------------------------------------------------------
// Body 2: TMA bulk-copy pipeline (cp.async.bulk, SASS UBLKCP) with mbarrier full/empty rings.
// Warp 0 is the producer: it issues one 1-D bulk copy per operand row per stage (BK doubles = 128 B
// into the padded shared row) and arms the stage's `full` barrier with the byte count; every warp
// waits on `full`, runs the stage's DMMAs, and one lane per warp arrives on `empty`.  No per-thread
// LDGSTS address arithmetic and no CTA-wide barrier in the main loop.
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <class C>
constexpr size_t bulk_smem_bytes() {
    return C::SMEM + 2 * C::STAGES * sizeof(uint64_t);
}

template <class C, bool WEIGHTED, bool PREFETCH>
__device__ __forceinline__ void gemm_tile_body_bulk(const KfGemmTask& t, double* smem) {
    constexpr int BK = C::BK, STAGES = C::STAGES, LDSP = C::LDSP, MI = C::MI, NJ = C::NJ;
    constexpr int NWARPS = C::THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / C::WN, wn = warp % C::WN;
    double* As = smem;
    double* Bs = smem + STAGES * C::A_STAGE;
    double* Ws = Bs + STAGES * C::B_STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(Ws + STAGES * BK);
    uint64_t* empty = full + STAGES;
    const int nk = (t.k1 - t.k0) / BK;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NWARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    constexpr unsigned ROW_BYTES = BK * sizeof(double);
    constexpr unsigned STAGE_BYTES = (C::BM + C::BN) * ROW_BYTES + (WEIGHTED ? ROW_BYTES : 0);
    auto produce = [&](int kt) {   // warp 0 only
        const int stage = kt % STAGES;
        const int k = t.k0 + kt * BK;
        if (lane == 0) mbar_expect_tx(full + stage, STAGE_BYTES);
        __syncwarp();
        double* as = As + stage * C::A_STAGE;
        double* bs = Bs + stage * C::B_STAGE;
        for (int r = lane; r < C::BM; r += 32)
            bulk_copy_g2s(as + r * LDSP, t.A + (long long)min(r, t.a_rows - 1) * t.lda + k, ROW_BYTES, full + stage);
        for (int r = lane; r < C::BN; r += 32)
            bulk_copy_g2s(bs + r * LDSP, t.B + (long long)min(r, t.b_rows - 1) * t.ldb + k, ROW_BYTES, full + stage);
        if (WEIGHTED && lane == 0) bulk_copy_g2s(Ws + stage * BK, t.W + k, ROW_BYTES, full + stage);
    };

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (warp == 0)
        for (int s = 0; s < STAGES - 1 && s < nk; ++s) produce(s);

    const int fr = (lane >> 2) * LDSP + (C::VEC ? 2 * (lane & 3) : (lane & 3));
    const int a_row0 = wm * (C::BM / C::WM), b_row0 = wn * (C::BN / C::WN);
    for (int kt = 0; kt < nk; ++kt) {
        const int stage = kt % STAGES;
        if (warp == 0) {
            // refill the slot consumed at iteration kt-1 with tile kt+STAGES-1 once every warp has released it
            const int nxt = kt + STAGES - 1;
            if (nxt < nk) {
                if (kt > 0) mbar_wait(empty + (nxt % STAGES), ((kt - 1) / STAGES) & 1);
                produce(nxt);
            }
        }
        mbar_wait(full + stage, (kt / STAGES) & 1);
        if (PREFETCH && t.accumulate && t.ldn == 1 && kt == nk - 4) prefetch_out_tile<C>(t, tid);
        mma_stage<C, WEIGHTED>(As + stage * C::A_STAGE + a_row0 * LDSP + fr, Bs + stage * C::B_STAGE + b_row0 * LDSP + fr,
                               Ws + stage * BK + (C::VEC ? 2 * (lane & 3) : (lane & 3)), acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);
    }
    store_tile<C>(t, acc, a_row0, b_row0, lane);
}

// ------------------------------------------------------------------------------------------------
// Body 3: tensor-map TMA (cp.async.bulk.tensor.2d, SASS UTMALDG) with SWIZZLE_128B.
// The panel is described once by a 2-D tensor map (inner dim = snapshots, outer = observable rows,
// box = 16 snapshots x 64 rows = 128-byte rows).  One elected thread issues 3 tensor copies + one
// 128-byte weight copy per stage; shared rows are dense 128 B and XOR-swizzled by the TMA unit
// (16-byte chunk c of row r lives at chunk c ^ (r & 7)).  To keep the LDS.128 fragment loads
// conflict-free under that swizzle the DMMA row index g = lane/4 is mapped to tile row
// (g >> 1) + 4 * (g & 1): the two rows of a quarter-warp then differ in bit 2 of (r & 7) and the
// eight 16-byte chunks they touch fill the 128-byte bank window exactly once.  The same permutation
// is applied to A rows, B rows and (inverted) in the epilogue.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

template <int STAGES_>
struct TmaCfg {
    static constexpr int BM = 128, BN = 64, BK = 16, STAGES = STAGES_, WM = 2, WN = 2, THREADS = 128, MINB = 2;
    static constexpr int MI = BM / WM / 8, NJ = BN / WN / 8;
    static constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128, W_BYTES = 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES + 1024;          // weights in their own 1 KB slot (keeps 1 KB alignment)
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;   // + alignment slack
};

template <class C, bool WEIGHTED>
__device__ __forceinline__ void gemm_tile_body_tma(const KfTmaTask& t, const void* tmap, unsigned char* smem_raw) {
    constexpr int STAGES = C::STAGES, MI = C::MI, NJ = C::NJ, BK = C::BK;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / C::WN, wn = warp % C::WN;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * C::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    const int nk = (t.k1 - t.k0) / BK;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, C::THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    constexpr unsigned TX_BYTES = C::A_BYTES + C::B_BYTES + (WEIGHTED ? C::W_BYTES : 0);
    auto produce = [&](int kt) {   // one elected thread
        unsigned char* st = smem + (size_t)(kt % STAGES) * C::STAGE_BYTES;
        uint64_t* bar = full + (kt % STAGES);
        const int k = t.k0 + kt * BK;
        mbar_expect_tx(bar, TX_BYTES);
        tma_load_2d(st, tmap, k, t.a_row, bar);                         // A rows   0..63
        tma_load_2d(st + 64 * 128, tmap, k, t.a_row + 64, bar);         // A rows  64..127
        tma_load_2d(st + C::A_BYTES, tmap, k, t.b_row, bar);            // B rows   0..63
        if (WEIGHTED) bulk_copy_g2s(st + C::A_BYTES + C::B_BYTES, t.W + k, C::W_BYTES, bar);
    };

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (tid == 0)
        for (int s = 0; s < STAGES - 1 && s < nk; ++s) produce(s);

    // fragment addressing under SWIZZLE_128B
    const int g = lane >> 2, q = lane & 3;
    const int rp = (g >> 1) + 4 * (g & 1);                 // permuted row inside an 8-row group; also (row & 7)
    const int a_row0 = wm * (C::BM / C::WM), b_row0 = wn * (C::BN / C::WN);
    const int a_off = (a_row0 + rp) * 128, b_off = C::A_BYTES + (b_row0 + rp) * 128;
    const int ch0 = ((0 * 4 + q) ^ rp) * 16, ch1 = ((1 * 4 + q) ^ rp) * 16;   // k2 = 0, 1

    for (int kt = 0; kt < nk; ++kt) {
        const int stage = kt % STAGES;
        if (tid == 0) {
            const int nxt = kt + STAGES - 1;
            if (nxt < nk) {
                if (kt > 0) mbar_wait(empty + (nxt % STAGES), ((kt - 1) / STAGES) & 1);
                produce(nxt);
            }
        }
        mbar_wait(full + stage, (kt / STAGES) & 1);
        if (t.out && kt == nk - 4) {   // pull the accumulator sub-tile into L2 ahead of the RMW epilogue
            const char* o = reinterpret_cast<const char*>(t.out);
            for (int l = tid; l < C::BM * 4; l += C::THREADS) prefetch_l2(o + (size_t)(l >> 2) * 128 * 8 + (size_t)(l & 3) * 128);
        }
        const unsigned char* st = smem + (size_t)stage * C::STAGE_BYTES;
        const double2 w0 = WEIGHTED ? *reinterpret_cast<const double2*>(st + C::A_BYTES + C::B_BYTES + (0 * 8 + 2 * q) * 8) : make_double2(1, 1);
        const double2 w1 = WEIGHTED ? *reinterpret_cast<const double2*>(st + C::A_BYTES + C::B_BYTES + (1 * 8 + 2 * q) * 8) : make_double2(1, 1);
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2) {
            const int ch = k2 ? ch1 : ch0;
            double2 a[MI], b[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const double2*>(st + a_off + i * 8 * 128 + ch);
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = *reinterpret_cast<const double2*>(st + b_off + j * 8 * 128 + ch);
            if (WEIGHTED) {
                const double2 w = k2 ? w1 : w0;
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
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);
    }
    // epilogue (accumulate into the 128-wide accumulator tile): DMMA (m = g, n = 2q + e) -> tile (row rp(g), col rp(2q+e))
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int m = a_row0 + i * 8 + rp;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            double* p = t.out + (size_t)m * 128 + b_row0 + j * 8;
            p[q] += acc[i][j][0];          // n index 2q   -> column q
            p[q + 4] += acc[i][j][1];      // n index 2q+1 -> column q + 4
        }
    }
}

}  // namespace kfg
