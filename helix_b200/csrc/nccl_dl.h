// NCCL entry points used by the replica weight broadcast (hb_model_load_broadcast), bound at first use with
// dlopen("libnccl.so.2"): a host process that already carries an NCCL (e.g. one that imported torch) keeps a single copy,
// and a single-GPU runner needs no NCCL installed at all.  Types come from the system <nccl.h>.
#pragma once
#include <nccl.h>

namespace hb {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*GetVersion)(int*);
};

// nullptr when libnccl.so.2 cannot be opened or lacks a symbol; *why then says which
const NcclApi* nccl_api(const char** why);

}  // namespace hb
