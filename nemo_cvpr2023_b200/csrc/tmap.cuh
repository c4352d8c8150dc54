// Host-side TMA tensor-map construction (cuTensorMapEncodeTiled fetched at run time: no libcuda link).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace nb {
// 2-D tensor [outer][inner] of `elem_bytes`-byte elements (4: fp32, 8: opaque 64-bit), row stride in bytes (a multiple
// of 16; base 16-byte aligned), box = box_inner x box_outer elements (each <= 256), no swizzle, out-of-bounds
// elements read as zero.  Results are cached per argument tuple.
int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t stride_bytes,
                 uint32_t box_inner, uint32_t box_outer);
}  // namespace nb
