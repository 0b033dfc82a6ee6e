#include "nccl_dl.h"

#include <dlfcn.h>

#include <mutex>
#include <string>

namespace hb {
namespace {
std::once_flag g_once;
NcclApi g_api{};
bool g_ok = false;
std::string g_why;

template <typename F>
bool bind(void* h, const char* name, F& out) {
  out = reinterpret_cast<F>(dlsym(h, name));
  if (!out) g_why = std::string("libnccl.so.2 lacks ") + name;
  return out != nullptr;
}
}  // namespace

const NcclApi* nccl_api(const char** why) {
  std::call_once(g_once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      const char* e = dlerror();
      g_why = std::string("cannot open libnccl.so.2: ") + (e ? e : "?");
      return;
    }
    g_ok = bind(h, "ncclGetUniqueId", g_api.GetUniqueId) && bind(h, "ncclCommInitRank", g_api.CommInitRank) &&
           bind(h, "ncclBroadcast", g_api.Broadcast) && bind(h, "ncclCommDestroy", g_api.CommDestroy) &&
           bind(h, "ncclGetErrorString", g_api.GetErrorString) && bind(h, "ncclGetVersion", g_api.GetVersion);
  });
  if (why) *why = g_why.c_str();
  return g_ok ? &g_api : nullptr;
}

}  // namespace hb
