// lip_comm.cuh — a NCCL communicator owned by the library (lip_comm of include/lip_b200.h).
// NCCL is resolved at run time (dlopen of the libnccl.so.2 the process already has — torch's bundled copy — else the system
// one), so the library itself carries no link-time dependency and loads on hosts without NCCL.
#pragma once
#include <nccl.h>

#include "lip_common.cuh"

struct lip_comm {
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
};

namespace lip {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};
// nullptr (and lip_last_error set) when NCCL cannot be loaded
const NcclApi* nccl_api();

#define LIP_CHECK_NCCL(api, expr)                                                                     \
  do {                                                                                                \
    ncclResult_t _r = (expr);                                                                         \
    if (_r != ncclSuccess) {                                                                          \
      ::lip::set_error("%s:%d NCCL error %s: %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r)); \
      return LIP_ERR_CUDA;                                                                            \
    }                                                                                                 \
  } while (0)
}  // namespace lip
