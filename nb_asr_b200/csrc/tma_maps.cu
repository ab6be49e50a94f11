// Host side of the TMA path: a cache of cuTensorMapEncodeTiled descriptors keyed by (base, dims, strides, box).
// Maps describe 2-byte elements (bf16 / fp16 alike: the copy engine never interprets them); rank 2 or 3,
// dims[0] contiguous, strides in ELEMENTS for dims 1, 2; swizzle: 1 = 128-byte (default), 0 = none, 2 = 64-byte.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "kernels.h"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

struct MapKey {
  uint64_t v[10];
  bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 10; ++i) { h ^= k.v[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

}  // namespace

int sm100_get_map(const void* base, int rank, const uint64_t* dims, const int64_t* strides_el, const uint32_t* box, CUtensorMap* out,
                  int swizzle128) {
  MapKey k{};
  k.v[0] = (uint64_t)base; k.v[1] = rank | ((swizzle128 ^ 1) << 4);
  for (int i = 0; i < rank; ++i) { k.v[2 + i] = dims[i]; k.v[7 + i] = box[i]; }
  for (int i = 1; i < rank; ++i) k.v[4 + i] = (uint64_t)strides_el[i];
  std::lock_guard<std::mutex> lk(g_maps_mu);
  auto it = g_maps.find(k);
  if (it != g_maps.end()) { *out = it->second; return 0; }
  if (g_maps.size() > 65536) g_maps.clear();     // plans come and go (Engine LRU): keep the cache bounded
  EncodeTiledFn enc = get_encode();
  if (!enc) return nbasr_fail("cuTensorMapEncodeTiled entry point not found");
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bx[3], es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 1; i < rank; ++i) gstr[i - 1] = (cuuint64_t)strides_el[i] * 2;
  if ((uint64_t)base % 16 != 0) return nbasr_fail("TMA base %p not 16-byte aligned", base);
  for (int i = 1; i < rank; ++i)
    if (gstr[i - 1] % 16 != 0) return nbasr_fail("TMA stride %llu not a multiple of 16 bytes", (unsigned long long)gstr[i - 1]);
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle128 == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return nbasr_fail("cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu %llu)", (int)r, rank,
                                           (unsigned long long)dims[0], (unsigned long long)dims[1],
                                           (unsigned long long)(rank > 2 ? dims[2] : 0));
  g_maps[k] = m;
  *out = m;
  return 0;
}
