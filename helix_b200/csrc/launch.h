// Launch helper: programmatic dependent launch (PDL) for the decode step's short kernels.
// A PDL-launched kernel may begin (prologue + anything that touches ONLY static data such as weights)
// while its predecessor drains; it must execute pdl_wait() before reading or writing any activation.
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace hb {

bool pdl_enabled();  // HB_PDL=0 disables (A/B measurements)

// SMs the calling thread's launches may occupy (0 = the whole device): hb_engine_cfg.sm_budget, set by the engine around
// its steps.  Every persistent / SM-count-sized grid (prefill GEMM workers, attention item walkers, stream-K decode GEMMs,
// split-KV heuristics) sizes itself to min(device SMs, limit) so engines packed on one GPU run side by side.
int& sm_limit();
int effective_sms(int device_sms);
struct SmLimitScope {  // RAII: engine entry points run on arbitrary caller threads
  int prev;
  explicit SmLimitScope(int lim) : prev(sm_limit()) { sm_limit() = lim; }
  ~SmLimitScope() { sm_limit() = prev; }
};
// Green-context stream: a stream whose kernels can only be scheduled on `sm_count` SMs (hardware partition).
// Returns false (with *why) when the driver cannot provide it.  `priority`: 0 default, 1 highest.
bool create_partition_stream(int device, int sm_count, int priority, cudaStream_t* stream, void** green_ctx, int* granted,
                             const char** why);
void destroy_partition(void* green_ctx);

template <typename... KArgs, typename... Args>
cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int n = 0;
  if (pdl && pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

}  // namespace hb
