// pf_comm.cu — the one collective of the data-parallel step behind the C ABI.
//
// Replaces mgw.DistributedOptimizer's per-variable Horovod all-reduces and mgw.broadcast_global_variables
// (/root/reference/utils/multi_gpu_wrapper.py:82-98; call sites learners/uniform_quantization/learner.py:245-247, :271)
// by ONE ncclAllReduce (sum) over the flat fp32 gradient buffer — or over a few contiguous buckets of it, each issued
// as soon as the backward pass has finished writing it — on a stream the caller chooses, so the transfer over
// NVLink / NVSwitch can be captured into the step's CUDA graph and overlapped with the rest of the backward pass.
// The division by the number of workers is folded into the optimizer kernels (grad_scale).
//
// NCCL is not linked: the library binds libnccl.so.2 at run time (the copy PyTorch ships is already mapped into the
// process; PF_NCCL_LIB names another path), so that single-GPU users never need it.  The communicator is created
// from a 128-byte unique id that rank 0 generates and the host side distributes (torch.distributed, MPI, a file ...).
#include <dlfcn.h>
#include <stdlib.h>

#include "pf_common.cuh"

namespace {
struct NcclUniqueId {
  char internal[128];
};
typedef void* NcclComm;
typedef int (*fn_get_uid)(NcclUniqueId*);
typedef int (*fn_init_rank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_destroy)(NcclComm);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_bcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_version)(int*);

struct Nccl {
  void* h = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_destroy destroy = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_bcast bcast = nullptr;
  fn_errstr errstr = nullptr;
  fn_version version = nullptr;
};
Nccl g_nccl;
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

int bind_nccl() {
  if (g_nccl.h) return PF_OK;
  const char* path = getenv("PF_NCCL_LIB");
  void* h = dlopen((path && *path) ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    pf_set_error("pf_comm: cannot load NCCL (%s); set PF_NCCL_LIB to libnccl.so.2", dlerror());
    return PF_ERR_NCCL;
  }
  Nccl n;
  n.h = h;
  n.get_uid = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
  n.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
  n.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
  n.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
  n.bcast = (fn_bcast)dlsym(h, "ncclBroadcast");
  n.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
  n.version = (fn_version)dlsym(h, "ncclGetVersion");
  if (!n.get_uid || !n.init_rank || !n.destroy || !n.allreduce || !n.bcast || !n.errstr) {
    pf_set_error("pf_comm: the NCCL library lacks a required symbol");
    dlclose(h);
    return PF_ERR_NCCL;
  }
  g_nccl = n;
  return PF_OK;
}

#define PF_NCCL(call, who)                                                          \
  do {                                                                              \
    const int r__ = (call);                                                         \
    if (r__ != 0) {                                                                 \
      pf_set_error("%s: NCCL error %d: %s", who, r__, g_nccl.errstr(r__));          \
      return PF_ERR_NCCL;                                                           \
    }                                                                               \
  } while (0)
}  // namespace

extern "C" {

int pf_comm_nccl_version(int* version_out) {
  PF_REQUIRE(version_out != nullptr, "pf_comm_nccl_version: null out");
  const int b = bind_nccl();
  if (b != PF_OK) return b;
  *version_out = 0;
  if (g_nccl.version) PF_NCCL(g_nccl.version(version_out), "pf_comm_nccl_version");
  return PF_OK;
}

int pf_comm_unique_id(void* id128_out) {
  PF_REQUIRE(id128_out != nullptr, "pf_comm_unique_id: null out");
  const int b = bind_nccl();
  if (b != PF_OK) return b;
  NcclUniqueId id;
  PF_NCCL(g_nccl.get_uid(&id), "pf_comm_unique_id");
  memcpy(id128_out, &id, sizeof(id));
  return PF_OK;
}

int pf_comm_init(const void* id128, int n_ranks, int rank, void** comm_out) {
  PF_REQUIRE(id128 && comm_out, "pf_comm_init: null pointer");
  PF_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "pf_comm_init: rank %d of %d", rank, n_ranks);
  const int b = bind_nccl();
  if (b != PF_OK) return b;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm c = nullptr;
  PF_NCCL(g_nccl.init_rank(&c, n_ranks, id, rank), "pf_comm_init");     // on the calling thread's current device
  *comm_out = c;
  return PF_OK;
}

int pf_comm_destroy(void* comm) {
  if (!comm) return PF_OK;
  PF_REQUIRE(g_nccl.h != nullptr, "pf_comm_destroy: NCCL is not loaded");
  PF_NCCL(g_nccl.destroy((NcclComm)comm), "pf_comm_destroy");
  return PF_OK;
}

int pf_allreduce_flat(void* comm, float* buf_dev, int64_t n, void* stream) {
  PF_REQUIRE(n >= 0, "pf_allreduce_flat: n < 0");
  if (n == 0) return PF_OK;
  PF_REQUIRE(comm && buf_dev, "pf_allreduce_flat: null pointer");
  PF_REQUIRE(g_nccl.h != nullptr, "pf_allreduce_flat: NCCL is not loaded (pf_comm_init first)");
  PF_NCCL(g_nccl.allreduce(buf_dev, buf_dev, (size_t)n, kNcclFloat32, kNcclSum, (NcclComm)comm, (cudaStream_t)stream),
          "pf_allreduce_flat");
  pf_count_launch();
  return PF_OK;
}

int pf_broadcast_flat(void* comm, float* buf_dev, int64_t n, int root, void* stream) {
  PF_REQUIRE(n >= 0 && root >= 0, "pf_broadcast_flat: bad arguments");
  if (n == 0) return PF_OK;
  PF_REQUIRE(comm && buf_dev, "pf_broadcast_flat: null pointer");
  PF_REQUIRE(g_nccl.h != nullptr, "pf_broadcast_flat: NCCL is not loaded (pf_comm_init first)");
  PF_NCCL(g_nccl.bcast(buf_dev, buf_dev, (size_t)n, kNcclFloat32, root, (NcclComm)comm, (cudaStream_t)stream),
          "pf_broadcast_flat");
  pf_count_launch();
  return PF_OK;
}

}  // extern "C"
