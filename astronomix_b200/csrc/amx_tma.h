// amx_tma.h -- tensor maps (TMA descriptors) over the [rows][x][y][z] arrays of the step.
//
// A CUtensorMap is built on the host with cuTensorMapEncodeTiled; the driver entry point is looked up
// through the runtime (cudaGetDriverEntryPoint), so libamx.so does not link against libcuda.  Maps are
// passed to kernels as __grid_constant__ parameters (the TMA unit reads them from constant memory).
// Device side: cp.async.bulk.tensor (SASS UTMALDG) into shared memory with mbarrier completion, and
// cp.async.bulk.prefetch.tensor (SASS UTMAPF) into L2.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "amx_internal.h"

namespace amx {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tma_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// fp32 map of `rows` rows of a View: dims (z, y, x, row).  xpad = resident ghost planes in front of the
// first active plane (slab ranks: 2): the map starts at the first ghost plane, so that they are in bounds
// and x coordinates are shifted by xpad.  Box = (bz, by, 1, brows).  Box coordinates may be negative and a
// box may stick out of the array (those elements read as zero / are skipped by a prefetch), but the z
// coordinate must be a multiple of 4: every box row has to start at a 16-byte aligned address, otherwise
// the instruction faults ("illegal instruction"; probed in tools/micro/tma_probe.cu).  Returns false when the layout cannot
// be described (extents / strides not multiples of 16 bytes) or the driver entry point is missing.
inline bool tma_map_rows_3d(CUtensorMap* out, const View& v, const Geom& g, int rows, int xpad, unsigned bz,
                            unsigned by, unsigned brows) {
  EncodeTiledFn enc = tma_encode_fn();
  if (!enc || g.nd != 3) return false;
  if (v.s[2] != 1 || (v.s[1] & 3) || (v.s[0] & 3) || (v.vs & 3) || (g.n[2] & 3)) return false;
  char* base = static_cast<char*>(v.p) - (size_t)xpad * (size_t)v.s[0] * sizeof(float);
  if (reinterpret_cast<uintptr_t>(base) & 15) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)g.n[2], (cuuint64_t)g.n[1], (cuuint64_t)(g.n[0] + 2 * xpad), (cuuint64_t)rows};
  const cuuint64_t strides[3] = {(cuuint64_t)v.s[1] * sizeof(float), (cuuint64_t)v.s[0] * sizeof(float),
                                 (cuuint64_t)v.vs * sizeof(float)};
  const cuuint32_t box[4] = {bz, by, 1, brows};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  if ((bz * sizeof(float)) & 15 || bz > 256 || by > 256 || brows > 256) return false;
  const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

#ifdef __CUDACC__
// pull the box at (z, y, x, row) into L2 (no shared memory, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* m, int z, int y, int x, int row) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(m), "r"(z), "r"(y),
               "r"(x), "r"(row)
               : "memory");
}
// box at (z, y, x, row) -> shared memory; completes `bytes` on the mbarrier
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, int z, int y, int x, int row,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"((unsigned)__cvta_generic_to_shared(smem_dst)),
      "l"(m), "r"(z), "r"(y), "r"(x), "r"(row), "r"((unsigned)__cvta_generic_to_shared(bar))
      : "memory");
}
#endif

}  // namespace amx
