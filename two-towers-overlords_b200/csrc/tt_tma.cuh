// Host-side TMA tensor-map construction.  cuTensorMapEncodeTiled is fetched through the runtime's driver
// entry-point query, so libtt_b200.so has no link-time dependency on libcuda (it must load, symbols only,
// on a machine without a driver).
#pragma once
#include <cuda.h>

#include "tt_common.cuh"

namespace tt {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// Row-major bf16 matrix [rows, cols] with row pitch `ld` elements, loaded as boxes of box_rows x 64 columns
// (64 bf16 = 128 B = one swizzle span) in the 128-byte-swizzled K-major layout tcgen05.mma reads.
// Out-of-range rows / columns are zero-filled, so neither extent needs to be a multiple of the box.
static inline int make_map_bf16_kmajor(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                                       uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  TT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  TT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0,
             "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ld=%llu)", (unsigned long long)ld);
  TT_REQUIRE(rows >= 1 && cols >= 1 && box_rows >= 1 && box_rows <= 256, "bad TMA extents");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace tt
