// Host-side construction of TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point, so
// libp3d.so does not link libcuda directly).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/p3d.h"

namespace p3d {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (EncodeTiledFn)p;
    }
    return fn;
}

// dims / box: innermost first; strides_bytes[i] = byte stride of dimension i+1. Out-of-bounds elements read as zero.
// elem_strides (optional): traversal stride per dimension; a dimension with stride s loads ceil(box / s) elements.
inline int make_tmap(CUtensorMap* tm, const void* base, CUtensorMapDataType dtype, CUtensorMapSwizzle swizzle, int rank,
                     const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides = nullptr) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return P3D_UNSUPPORTED;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(tm, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? P3D_OK : P3D_BAD_ARG;
}

}  // namespace p3d
