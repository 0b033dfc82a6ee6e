// Host-side construction of TMA tensor maps (cuTensorMapEncodeTiled resolved through the runtime's
// driver entry point, so the library links against libcudart only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

enum TmDtype { TM_BF16 = 0, TM_F32 = 1 };

// 2-D row-major tensor [outer][inner]; box = [box_outer][box_inner]; 128B swizzle when the box row is
// exactly 128 bytes, no swizzle otherwise. Out-of-bounds elements read as zero / are not written.
bool make_tmap_2d(CUtensorMap* out, const void* gptr, TmDtype dt, uint64_t inner, uint64_t outer,
                  uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);
// 3-D tensor [d2][d1][d0] with byte strides for d1, d2.
bool make_tmap_3d(CUtensorMap* out, const void* gptr, TmDtype dt, uint64_t d0, uint64_t d1, uint64_t d2,
                  uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2);
const char* tmap_last_error();

}  // namespace hb
