// Host helper: encode the 2-D tensor map of a lifted panel (rows x Mc doubles, row-major with the snapshot
// index contiguous) for cp.async.bulk.tensor with SWIZZLE_128B, box = 16 snapshots x 64 rows.
// cuTensorMapEncodeTiled is fetched through the runtime (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*kf_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline int kf_make_panel_tensor_map(CUtensorMap* out, void* base, unsigned long long Mc, unsigned long long rows) {
    static kf_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return -1;
        fn = reinterpret_cast<kf_encode_tiled_fn>(p);
    }
    const cuuint64_t gdim[2] = {Mc, rows};
    const cuuint64_t gstride[1] = {Mc * sizeof(double)};   // bytes between consecutive rows
    const cuuint32_t box[2] = {16, 64};                    // 16 doubles = 128 B inner, 64 rows
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)r;
}
