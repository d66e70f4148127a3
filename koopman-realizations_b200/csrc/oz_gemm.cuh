// INT8 tensor-core contraction on the 5th-generation tensor cores (tcgen05.mma kind::i8, INT32 accumulators in TMEM) — the
// inner product engine of the FP64-exact "Ozaki scheme II" Gram path (ozaki.cu):
//
//     R[m][n] = ( sum_k A[m][k] * B[n][k] )  mod p            A, B: int8 residues (K-major, k = snapshot), R: uint8 in [0, p)
//
// One persistent CTA per SM, warp-specialised (blackwell_cuda_programming.md, "canonical GEMM kernel"):
//   warp 0   TMA producer   cp.async.bulk.tensor.3d (k, row, modulus) -> 4-stage shared-memory ring, SWIZZLE_128B,
//                           A box 128 k x 128 rows (16 KB), B box 128 k x 256 rows (32 KB), mbarrier full / empty
//   warp 1   MMA issuer     one elected thread: 4 x tcgen05.mma.cta_group::1.kind::i8 (M 128, N 256, K 32) per stage from
//                           shared-memory descriptors; tcgen05.commit frees the stage and, after the last k block, hands the
//                           accumulator (one of two 256-column TMEM buffers) to the epilogue
//   warps 2-5 epilogue      tcgen05.ld 32x32b (lane = output row) -> reduction mod p by a multiply-high -> uint8 rows to global;
//                           overlaps the next tile's main loop through the second TMEM buffer
// A task list (tile coordinates + modulus index) is walked round-robin, so one launch covers every tile of every modulus.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace oz {

constexpr int BM = 128, BN = 256, BK = 128;          // BK in bytes = int8 elements
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct Task {
    int a_row;        // first row of the A operand (tile of BM rows)
    int b_row;        // first row of the B operand (tile of BN rows)
    int t;            // modulus index (third tensor-map coordinate)
    int b_is_y;       // 0: B rows come from the X residues (Gram), 1: from the Y residues (cross product)
    unsigned long long out_off;   // element offset of out[a_row][b_row] inside the modulus-t output plane
};

struct Params {
    const Task* tasks;
    int ntasks;
    int K;                      // contraction length (multiple of BK)
    uint8_t* out;               // [T][planes...] uint8 residues
    unsigned long long plane;   // elements per modulus plane of `out`
    int ld_out;                 // row stride of the output planes (elements)
    int m_valid, n_valid_x, n_valid_y;   // rows beyond these are padding (not stored)
    unsigned p[32];             // moduli
    unsigned long long magic[32];   // ceil(2^37 / p): floor(a / p) = (a * magic) >> 37 exactly for a < 2^29
    int offset[32];             // multiple of p >= 2^27 (|accumulator| <= K * 128^2 <= 2^27 for K <= 8192): makes it non-negative
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OZ_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OZ_DONE;\n"
        "bra OZ_WAIT;\n"
        "OZ_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B (8-row x 128-byte atoms, 1024 bytes apart), sm_100 descriptor version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                      // leading byte offset: unused for swizzled K-major layouts
    d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = S32, A = B = signed 8 bit, both K-major, N = 256, M = 128
__device__ __forceinline__ constexpr uint32_t make_idesc() {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// The kernel.  tmap_xa: X residues with a 128-row box; tmap_xb / tmap_yb: X / Y residues with a 256-row box.
__global__ void __launch_bounds__(THREADS, 1)
oz_gemm_kernel(const __grid_constant__ CUtensorMap tmap_xa, const __grid_constant__ CUtensorMap tmap_xb,
               const __grid_constant__ CUtensorMap tmap_yb, const __grid_constant__ Params prm) {
    extern __shared__ __align__(1024) unsigned char oz_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;      // [2] accumulator buffer ready for the epilogue
    uint64_t* tempty = tfull + 2;          // [2] accumulator buffer drained
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = prm.K / BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull + b, 1); mbar_init(tempty + b, 4); }    // 4 epilogue warps arrive
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_base_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int it = 0;                                      // global k-block counter of this CTA
            for (int task = blockIdx.x; task < prm.ntasks; task += gridDim.x) {
                const Task tk = prm.tasks[task];
                const CUtensorMap* bmap = tk.b_is_y ? &tmap_yb : &tmap_xb;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % STAGES;
                    if (it >= STAGES) mbar_wait(empty + s, ((it / STAGES) - 1) & 1);
                    unsigned char* st = smem + (size_t)s * STAGE_BYTES;
                    mbar_expect_tx(full + s, STAGE_BYTES);
                    tma_load_3d(st, &tmap_xa, kb * BK, tk.a_row, tk.t, full + s);
                    tma_load_3d(st + A_BYTES, bmap, kb * BK, tk.b_row, tk.t, full + s);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc();
            int it = 0, tile = 0;
            for (int task = blockIdx.x; task < prm.ntasks; task += gridDim.x, ++tile) {
                const int buf = tile & 1;
                if (tile >= 2) mbar_wait(tempty + buf, ((tile >> 1) - 1) & 1);       // epilogue has drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(full + s, (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES);
                    const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 32; ++k)      // K = 32 int8 per instruction: +32 bytes = +2 in the 16-byte address field
                        mma_i8(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
                    mma_commit(empty + s);                 // the stage is free once these MMAs have read it
                }
                mma_commit(tfull + buf);                   // accumulator complete
            }
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane group = warp % 4 =====
        const int lg = warp & 3;
        int tile = 0;
        for (int task = blockIdx.x; task < prm.ntasks; task += gridDim.x, ++tile) {
            const Task tk = prm.tasks[task];
            const int buf = tile & 1;
            mbar_wait(tfull + buf, (tile >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const unsigned p = prm.p[tk.t];
            const unsigned long long mg = prm.magic[tk.t];
            const int off = prm.offset[tk.t];
            const int row = tk.a_row + lg * 32 + lane;
            const int n_valid = tk.b_is_y ? prm.n_valid_y : prm.n_valid_x;
            uint8_t* orow = prm.out + (size_t)tk.t * prm.plane + tk.out_off + (size_t)(lg * 32 + lane) * prm.ld_out;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * BN + c0), v);
                uint32_t packed[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    uint32_t w = 0;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const unsigned a = (unsigned)((int)v[q * 4 + e] + off);              // >= 0, < 2^28 + p
                        const unsigned quo = (unsigned)(((unsigned long long)a * mg) >> 37);
                        const unsigned r = a - quo * p;                                        // in [0, p)
                        w |= (r & 0xFFu) << (8 * e);
                    }
                    packed[q] = w;
                }
                if (row < prm.m_valid && tk.b_row + c0 < n_valid) {
                    uint4* dst = reinterpret_cast<uint4*>(orow + c0);
                    dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty + buf);
        }
    }
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(TMEM_COLS));
}

}  // namespace oz
