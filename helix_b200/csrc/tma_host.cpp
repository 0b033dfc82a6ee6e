#include "tma_host.h"

#include <stdlib.h>
#include <string.h>

#include "launch.h"

#include <stdio.h>

#include <mutex>

namespace hb {
namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;
thread_local char g_err[256] = "";

void resolve() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<EncodeTiledFn>(fn);
}
}  // namespace

const char* tmap_last_error() { return g_err; }

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("HB_PDL");
    return !(e && !strcmp(e, "0"));
  }();
  return on;
}

bool make_tmap_2d(CUtensorMap* out, const void* gptr, TmDtype dt, uint64_t inner, uint64_t outer,
                  uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  std::call_once(g_once, resolve);
  if (!g_encode) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return false;
  }
  const uint32_t esz = dt == TM_BF16 ? 2 : 4;
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = (box_inner * esz == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(out, dt == TM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<void*>(gptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled(2d) failed: %d (ptr=%p inner=%llu outer=%llu stride=%llu box=%ux%u)",
             (int)r, gptr, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes,
             box_inner, box_outer);
    return false;
  }
  return true;
}

bool make_tmap_3d(CUtensorMap* out, const void* gptr, TmDtype dt, uint64_t d0, uint64_t d1, uint64_t d2,
                  uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  std::call_once(g_once, resolve);
  if (!g_encode) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return false;
  }
  const uint32_t esz = dt == TM_BF16 ? 2 : 4;
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = (b0 * esz == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(out, dt == TM_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                        const_cast<void*>(gptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
    return false;
  }
  return true;
}

}  // namespace hb
