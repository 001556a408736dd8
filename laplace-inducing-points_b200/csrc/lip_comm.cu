// lip_comm.cu — library-owned NCCL communicators for the sharded Krylov recurrences (the reference is single-device,
// src/data.py:90-93).  One process per GPU; the caller exchanges the 128-byte unique id over whatever it has (torch.distributed
// in the Python mirror) and every rank of the group calls lip_comm_create.
#include <dlfcn.h>

#include <mutex>

#include "lip_comm.cuh"

namespace lip {

static NcclApi g_api;
static bool g_api_ok = false;
static std::once_flag g_api_once;

static void load_api() {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy this process already uses (torch's), if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW);
  if (!h) return;
  bool ok = true;
  auto sym = [&](const char* name) {
    void* p = dlsym(h, name);
    if (!p) ok = false;
    return p;
  };
  g_api.GetUniqueId = (decltype(g_api.GetUniqueId))sym("ncclGetUniqueId");
  g_api.CommInitRank = (decltype(g_api.CommInitRank))sym("ncclCommInitRank");
  g_api.CommDestroy = (decltype(g_api.CommDestroy))sym("ncclCommDestroy");
  g_api.GetErrorString = (decltype(g_api.GetErrorString))sym("ncclGetErrorString");
  g_api.AllReduce = (decltype(g_api.AllReduce))sym("ncclAllReduce");
  g_api.AllGather = (decltype(g_api.AllGather))sym("ncclAllGather");
  g_api.ReduceScatter = (decltype(g_api.ReduceScatter))sym("ncclReduceScatter");
  g_api.GroupStart = (decltype(g_api.GroupStart))sym("ncclGroupStart");
  g_api.GroupEnd = (decltype(g_api.GroupEnd))sym("ncclGroupEnd");
  g_api_ok = ok;
}

const NcclApi* nccl_api() {
  std::call_once(g_api_once, load_api);
  if (!g_api_ok) {
    set_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "symbols missing");
    return nullptr;
  }
  return &g_api;
}

}  // namespace lip

using namespace lip;

extern "C" {

int lip_comm_unique_id(void* id128) {
  LIP_REQUIRE(id128, "lip_comm_unique_id: null buffer");
  const NcclApi* api = nccl_api();
  if (!api) return LIP_ERR_UNSUPPORTED;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  LIP_CHECK_NCCL(api, api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return LIP_OK;
}

int lip_comm_create(const void* id128, int32_t world, int32_t rank, lip_comm** out) {
  LIP_REQUIRE(id128 && out && world >= 1 && rank >= 0 && rank < world, "lip_comm_create: bad argument");
  const NcclApi* api = nccl_api();
  if (!api) return LIP_ERR_UNSUPPORTED;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  lip_comm* c = new lip_comm();
  c->world = world;
  c->rank = rank;
  ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank: %s", api->GetErrorString(r));
    delete c;
    return LIP_ERR_CUDA;
  }
  *out = c;
  return LIP_OK;
}

int lip_comm_destroy(lip_comm* c) {
  if (!c) return LIP_OK;
  const NcclApi* api = nccl_api();
  if (api && c->comm) api->CommDestroy(c->comm);
  delete c;
  return LIP_OK;
}

int lip_comm_world(const lip_comm* c) { return c ? c->world : 1; }
int lip_comm_rank(const lip_comm* c) { return c ? c->rank : 0; }

// in-place sum over the ranks of the communicator (tests / the scalar exchanges of the Python mirror)
int lip_comm_allreduce_sum(lip_comm* c, float* buf, int64_t count, lip_stream_t stream) {
  LIP_REQUIRE(c && buf && count > 0, "lip_comm_allreduce_sum: bad argument");
  const NcclApi* api = nccl_api();
  if (!api) return LIP_ERR_UNSUPPORTED;
  LIP_CHECK_NCCL(api, api->AllReduce(buf, buf, (size_t)count, ncclFloat, ncclSum, c->comm, (cudaStream_t)stream));
  return LIP_OK;
}

}  // extern "C"
