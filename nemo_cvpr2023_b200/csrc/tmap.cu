#include <array>
#include <map>
#include <mutex>
#include "common.cuh"
#include "tmap.cuh"

namespace nb {
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}
}  // namespace

int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t stride_bytes,
                 uint32_t box_inner, uint32_t box_outer) {
  static std::mutex mu;
  static std::map<std::array<uint64_t, 7>, CUtensorMap> cache;
  const std::array<uint64_t, 7> key = {(uint64_t)(uintptr_t)ptr, (uint64_t)elem_bytes, inner, outer, stride_bytes,
                                       box_inner, box_outer};
  std::lock_guard<std::mutex> lk(mu);
  auto f = cache.find(key);
  if (f != cache.end()) {
    *out = f->second;
    return 0;
  }
  EncodeTiledFn enc = encode_fn();
  NB_REQUIRE(enc, "make_tmap_2d: cuTensorMapEncodeTiled entry point unavailable");
  NB_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "make_tmap_2d: element size %d", elem_bytes);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("make_tmap_2d: cuTensorMapEncodeTiled failed (%d) for %llu x %llu x %d B, stride %llu, box %u x %u", (int)r,
              (unsigned long long)outer, (unsigned long long)inner, elem_bytes, (unsigned long long)stride_bytes,
              box_inner, box_outer);
    return NEMO_B200_ECUDA;
  }
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return 0;
}
}  // namespace nb
