// stats_allreduce.cu -- the ONE exchange step of the path behind the C ABI: the int64 batch moments of the per-sample bin counts (what np.mean / np.std over the
// batch axis need, analysis/eval_eig.py:620-623, :677-680) summed over the GPUs that each hold a slice of the analysis batch.
//
// NCCL is resolved at run time (dlopen: the copy already in the process -- e.g. the one PyTorch loaded -- else EIGB200_NCCL_LIB, else libnccl.so.2), so that
// libeigb200.so itself has no link-time dependency on it.  The caller brings the communicator: any ncclComm_t (one process per GPU, as torch.distributed /
// the reference's launcher would create it), or -- single process driving several GPUs -- eigb200_stats_comm_init_all.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

namespace eigb200 {
namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi& api() {
  static NcclApi a;
  static bool tried = false;
  if (tried) return a;
  tried = true;
  const char* env = getenv("EIGB200_NCCL_LIB");
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { a.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (a.handle) break; }     // a copy already loaded by the host framework
  if (!a.handle && env) a.handle = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  for (const char* n : names) { if (a.handle) break; a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); }
  if (!a.handle) return a;
#define EIGB_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name))
  EIGB_SYM(AllReduce, "ncclAllReduce"); EIGB_SYM(CommInitAll, "ncclCommInitAll"); EIGB_SYM(CommDestroy, "ncclCommDestroy"); EIGB_SYM(CommCount, "ncclCommCount");
  EIGB_SYM(GroupStart, "ncclGroupStart"); EIGB_SYM(GroupEnd, "ncclGroupEnd"); EIGB_SYM(GetErrorString, "ncclGetErrorString");
#undef EIGB_SYM
  a.ok = a.AllReduce && a.CommInitAll && a.CommDestroy && a.GroupStart && a.GroupEnd && a.GetErrorString;
  return a;
}
int nccl_fail(ncclResult_t r, const char* what) {
  set_error("%s failed: %s", what, api().GetErrorString ? api().GetErrorString(r) : "NCCL error");
  return EIGB200_ECUDA;
}
}  // namespace
}  // namespace eigb200

using namespace eigb200;

#define EIGB_NCCL_READY()                                                                                                        \
  do { if (!api().ok) { set_error("NCCL is not available (tried the loaded libnccl.so.2, EIGB200_NCCL_LIB, libnccl.so.2)"); return EIGB200_EUNSUPPORTED; } } while (0)

extern "C" int eigb200_stats_available(void) { return api().ok ? 1 : 0; }

extern "C" int eigb200_stats_comm_init_all(int ndev, const int* devs, void** comms) {
  EIGB_CHECK_ARG(ndev > 0 && comms, "stats_comm_init_all: bad arguments");
  EIGB_NCCL_READY();
  ncclResult_t r = api().CommInitAll(reinterpret_cast<ncclComm_t*>(comms), ndev, devs);
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitAll");
  return EIGB200_OK;
}

extern "C" int eigb200_stats_comm_destroy(void* comm) {
  EIGB_CHECK_ARG(comm, "stats_comm_destroy: null communicator");
  EIGB_NCCL_READY();
  ncclResult_t r = api().CommDestroy(reinterpret_cast<ncclComm_t>(comm));
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommDestroy");
  return EIGB200_OK;
}

extern "C" int eigb200_stats_group_start(void) { EIGB_NCCL_READY(); ncclResult_t r = api().GroupStart(); return r == ncclSuccess ? EIGB200_OK : nccl_fail(r, "ncclGroupStart"); }
extern "C" int eigb200_stats_group_end(void) { EIGB_NCCL_READY(); ncclResult_t r = api().GroupEnd(); return r == ncclSuccess ? EIGB200_OK : nccl_fail(r, "ncclGroupEnd"); }

extern "C" int eigb200_stats_allreduce(void* stream, void* comm, int64_t* d_moments, size_t count) {
  EIGB_CHECK_ARG(comm && d_moments && count > 0, "stats_allreduce: bad arguments");
  EIGB_NCCL_READY();
  ncclResult_t r = api().AllReduce(d_moments, d_moments, count, ncclInt64, ncclSum, reinterpret_cast<ncclComm_t>(comm), (cudaStream_t)stream);
  if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
  return EIGB200_OK;
}
