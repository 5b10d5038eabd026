// The pipeline's one data-path collective behind the C ABI: the all-gather of the final
// [n_local, K*D] descriptors before retrieval (BASELINE config 4; the reference has no multi-GPU path -- its
// get_top_k_recall, /root/reference/utilities.py:433-450, sees one host's database).  The library does not link NCCL:
// the host process already has it loaded (PyTorch's `torch.distributed` NCCL backend owns the communicator), so the
// symbol is resolved from the loaded libnccl at first use and the communicator comes in as an opaque ncclComm_t.
#include <dlfcn.h>
#include "common.cuh"

namespace {
typedef int (*AllGatherFn)(const void* sendbuff, void* recvbuff, size_t sendcount, int datatype, void* comm, cudaStream_t stream);
typedef const char* (*ErrStrFn)(int);
constexpr int kNcclFloat32 = 7;          // ncclDataType_t::ncclFloat32 (nccl.h; stable across NCCL 2.x)

void* nccl_symbol(const char* name) {
  void* s = dlsym(RTLD_DEFAULT, name);
  if (s) return s;
  for (const char* lib : {"libnccl.so.2", "libnccl.so"}) {
    void* h = dlopen(lib, RTLD_NOW | RTLD_NOLOAD);        // only a library the process has ALREADY loaded
    if (h && (s = dlsym(h, name))) return s;
  }
  return nullptr;
}
}  // namespace

extern "C" int anyloc_allgather_desc(void* nccl_comm, const float* local, float* all, size_t n_loc, int Dv, void* stream) {
  ANYLOC_REQUIRE(nccl_comm && local && all && Dv > 0, "allgather_desc: null communicator / pointer or bad Dv=%d", Dv);
  static AllGatherFn fn = nullptr;
  static ErrStrFn errstr = nullptr;
  if (!fn) {
    fn = (AllGatherFn)nccl_symbol("ncclAllGather");
    errstr = (ErrStrFn)nccl_symbol("ncclGetErrorString");
  }
  if (!fn) {
    anyloc::set_error("allgather_desc: NCCL is not loaded in this process (the caller owns the communicator and the library)");
    return ANYLOC_ERR_UNSUPPORTED;
  }
  if (n_loc == 0) return ANYLOC_OK;
  const int rc = fn(local, all, n_loc * (size_t)Dv, kNcclFloat32, nccl_comm, (cudaStream_t)stream);
  if (rc != 0) {
    anyloc::set_error("allgather_desc: ncclAllGather failed (%d: %s)", rc, errstr ? errstr(rc) : "?");
    return ANYLOC_ERR_CUDA;
  }
  anyloc::count_launch();
  return ANYLOC_OK;
}
