// Host-side construction of TMA tensor maps (CUtensorMap).  The encoder is a driver
// entry point; it is resolved through the runtime (cudaGetDriverEntryPoint), so the
// library still has no link-time dependency on libcuda.
#pragma once
#include <cuda.h>

#include "common.h"

namespace b200 {

typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                           CUtensorMapFloatOOBfill);

inline TensorMapEncodeTiledFn tensor_map_encoder() {
    static TensorMapEncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<TensorMapEncodeTiledFn>(p);
    }();
    return fn;
}

// rank-`rank` tiled map over raw bytes-like elements of `elem_bytes` (1, 2, 4 or 8).
// dims/box are in elements, innermost first; strides (bytes) are for dims 1..rank-1.
inline int make_tensor_map(CUtensorMap* out, int elem_bytes, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return fail(B200_E_NOLIB, "cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMapDataType dt;
    switch (elem_bytes) {
        case 1: dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
        case 2: dt = CU_TENSOR_MAP_DATA_TYPE_UINT16; break;
        case 4: dt = CU_TENSOR_MAP_DATA_TYPE_UINT32; break;
        case 8: dt = CU_TENSOR_MAP_DATA_TYPE_UINT64; break;
        default: return fail(B200_E_INVALID, "tensor map: element size %d", elem_bytes);
    }
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
    }
    CUresult r = enc(out, dt, cuuint32_t(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(int(r), "cuTensorMapEncodeTiled failed (CUresult %d)", int(r));
    return 0;
}

}  // namespace b200
