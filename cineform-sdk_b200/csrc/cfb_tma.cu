// cfb_tma.cu -- host side of the TMA path: tensor-map encoding through the driver entry point (cudart is linked
// statically and libcuda is never linked; cuTensorMapEncodeTiled is resolved with cudaGetDriverEntryPoint).
#include "cfb_tma.cuh"

#include <mutex>

namespace cfb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    });
    return fn;
}

cudaError_t tmap_encode_2d(CUtensorMap *out, const void *base, uint64_t row_bytes, uint64_t rows, uint64_t pitch_bytes,
                           uint32_t box_bytes, uint32_t box_rows, int elem_bytes)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    if (elem_bytes != 4 && elem_bytes != 8) return cudaErrorInvalidValue;
    const cuuint64_t dims[2] = {row_bytes / elem_bytes, rows};
    const cuuint64_t strides[1] = {pitch_bytes};
    const cuuint32_t box[2] = {box_bytes / elem_bytes, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(out, elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t tmap_encode_3d(CUtensorMap *out, const void *base, uint64_t row_bytes, uint64_t rows, uint64_t pitch_bytes,
                           uint64_t planes, uint64_t plane_bytes, uint32_t box_bytes, uint32_t box_rows, uint32_t box_planes)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    const cuuint64_t dims[3] = {row_bytes / 4, rows, planes};
    const cuuint64_t strides[2] = {pitch_bytes, plane_bytes};
    const cuuint32_t box[3] = {box_bytes / 4, box_rows, box_planes};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace cfb
