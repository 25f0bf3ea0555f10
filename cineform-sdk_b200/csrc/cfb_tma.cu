// cfb_tma.cu -- host side of the TMA path: tensor-map encoding through the driver entry point (cudart is linked
// statically and libcuda is never linked; cuTensorMapEncodeTiled is resolved with cudaGetDriverEntryPoint).
#include "cfb_tma.cuh"

#include <cstring>
#include <mutex>
#include <unordered_map>

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

static cudaError_t tmap_encode_2d_uncached(CUtensorMap *out, const void *base, uint64_t row_bytes, uint64_t rows, uint64_t pitch_bytes,
                                           uint32_t box_bytes, uint32_t box_rows, int elem_bytes);
static cudaError_t tmap_encode_3d_uncached(CUtensorMap *out, const void *base, uint64_t row_bytes, uint64_t rows, uint64_t pitch_bytes,
                                           uint64_t planes, uint64_t plane_bytes, uint32_t box_bytes, uint32_t box_rows, uint32_t box_planes);

// Encoded maps are memoised per host thread: a codec presents the same (base, geometry, box) tuples launch after launch
// (its device staging never moves), and one cuTensorMapEncodeTiled costs about a microsecond of host time -- the final
// inverse level alone needs six maps per frame of a batch.
namespace {
struct MapKey {
    uint64_t v[9];
    bool operator==(const MapKey &o) const { return !memcmp(v, o.v, sizeof(v)); }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const {
        uint64_t h = 1469598103934665603ull;
        for (uint64_t x : k.v) { h ^= x; h *= 1099511628211ull; h ^= h >> 29; }
        return (size_t)h;
    }
};
typedef std::unordered_map<MapKey, CUtensorMap, MapKeyHash> MapCache;
MapCache &map_cache() { static thread_local MapCache c; return c; }
bool cache_get(const MapKey &k, CUtensorMap *out)
{
    MapCache &c = map_cache();
    auto it = c.find(k);
    if (it == c.end()) return false;
    *out = it->second;
    return true;
}
void cache_put(const MapKey &k, const CUtensorMap &m)
{
    MapCache &c = map_cache();
    if (c.size() >= 8192) c.clear();
    c.emplace(k, m);
}
}  // namespace

cudaError_t tmap_encode_2d(CUtensorMap *out, const void *base, uint64_t row_bytes, uint64_t rows, uint64_t pitch_bytes,
                           uint32_t box_bytes, uint32_t box_rows, int elem_bytes)
{
    const MapKey key = {{(uint64_t)(uintptr_t)base, row_bytes, rows, pitch_bytes, 1, 0, box_bytes, ((uint64_t)box_rows << 32) | 1u, (uint64_t)elem_bytes}};
    if (cache_get(key, out)) return cudaSuccess;
    const cudaError_t e = tmap_encode_2d_uncached(out, base, row_bytes, rows, pitch_bytes, box_bytes, box_rows, elem_bytes);
    if (e == cudaSuccess) cache_put(key, *out);
    return e;
}

cudaError_t tmap_encode_3d(CUtensorMap *out, const void *base, uint64_t row_bytes, uint64_t rows, uint64_t pitch_bytes,
                           uint64_t planes, uint64_t plane_bytes, uint32_t box_bytes, uint32_t box_rows, uint32_t box_planes)
{
    const MapKey key = {{(uint64_t)(uintptr_t)base, row_bytes, rows, pitch_bytes, planes, plane_bytes, box_bytes, ((uint64_t)box_rows << 32) | box_planes, 4}};
    if (cache_get(key, out)) return cudaSuccess;
    const cudaError_t e = tmap_encode_3d_uncached(out, base, row_bytes, rows, pitch_bytes, planes, plane_bytes, box_bytes, box_rows, box_planes);
    if (e == cudaSuccess) cache_put(key, *out);
    return e;
}

static cudaError_t tmap_encode_2d_uncached(CUtensorMap *out, const void *base, uint64_t row_bytes, uint64_t rows, uint64_t pitch_bytes,
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

static cudaError_t tmap_encode_3d_uncached(CUtensorMap *out, const void *base, uint64_t row_bytes, uint64_t rows, uint64_t pitch_bytes,
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
