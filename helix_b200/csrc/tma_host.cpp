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

int& sm_limit() {
  thread_local int lim = 0;
  return lim;
}
int effective_sms(int device_sms) {
  const int lim = sm_limit();
  return lim > 0 ? (lim < device_sms ? lim : device_sms) : device_sms;
}

// ---- green contexts through the runtime's driver entry points (the library links libcudart only)
namespace {
template <typename F>
bool drv(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return false;
  fn = reinterpret_cast<F>(p);
  return true;
}
}  // namespace

bool create_partition_stream(int device, int sm_count, int priority, cudaStream_t* stream, void** green_ctx, int* granted,
                             const char** why) {
  static thread_local char msg[160];
  auto fail = [&](const char* what, int code) {
    snprintf(msg, sizeof msg, "green context: %s failed (%d)", what, code);
    if (why) *why = msg;
    return false;
  };
  CUresult (*getRes)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
  CUresult (*split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
  CUresult (*genDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
  CUresult (*gcCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
  CUresult (*gcStream)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
  CUresult (*devGet)(CUdevice*, int) = nullptr;
  if (!drv("cuDeviceGetDevResource", getRes) || !drv("cuDevSmResourceSplitByCount", split) ||
      !drv("cuDevResourceGenerateDesc", genDesc) || !drv("cuGreenCtxCreate", gcCreate) ||
      !drv("cuGreenCtxStreamCreate", gcStream) || !drv("cuDeviceGet", devGet))
    return fail("driver entry points", -1);
  CUdevice dev;
  CUresult r = devGet(&dev, device);
  if (r != CUDA_SUCCESS) return fail("cuDeviceGet", (int)r);
  CUdevResource all, part, rest;
  if ((r = getRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM)) != CUDA_SUCCESS) return fail("cuDeviceGetDevResource", (int)r);
  unsigned int groups = 1;
  if ((r = split(&part, &groups, &all, &rest, 0, (unsigned int)sm_count)) != CUDA_SUCCESS || groups < 1)
    return fail("cuDevSmResourceSplitByCount", (int)r);
  CUdevResourceDesc desc;
  if ((r = genDesc(&desc, &part, 1)) != CUDA_SUCCESS) return fail("cuDevResourceGenerateDesc", (int)r);
  CUgreenCtx gc;
  if ((r = gcCreate(&gc, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM)) != CUDA_SUCCESS) return fail("cuGreenCtxCreate", (int)r);
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = least urgent (numerically largest), hi = most urgent
  CUstream st;
  if ((r = gcStream(&st, gc, CU_STREAM_NON_BLOCKING, priority > 0 ? hi : lo)) != CUDA_SUCCESS) {
    destroy_partition(gc);
    return fail("cuGreenCtxStreamCreate", (int)r);
  }
  *stream = reinterpret_cast<cudaStream_t>(st);
  *green_ctx = gc;
  if (granted) *granted = (int)part.sm.smCount;
  return true;
}

void destroy_partition(void* green_ctx) {
  CUresult (*gcDestroy)(CUgreenCtx) = nullptr;
  if (green_ctx && drv("cuGreenCtxDestroy", gcDestroy)) gcDestroy(reinterpret_cast<CUgreenCtx>(green_ctx));
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
