// Row-sharded iteration loop with the all-reduce issued from C on the compute stream.
//
// rank r owns a block of rows of C; every iteration needs one all-reduce (sum) of
// m fp32 column partials.  Driving that loop from Python costs five enqueues per iteration through two
// libraries and two streams; at 8 GPUs an iteration is ~0.4 ms of device time and the host becomes the
// bottleneck.  Here the whole loop -- sweep, partial fold, ncclAllReduce, finalize -- is queued by one C
// call on ONE stream, so it is also CUDA-graph capturable.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already mapped into the process by PyTorch), so the
// library has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace b200ot {

struct Id128 {
  char bytes[128];  // same size and by-value ABI as ncclUniqueId
};

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
static NcclApi& nccl() {
  static NcclApi api;
  if (api.handle) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return api;
  api.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(api.handle, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(void**, int, Id128, int)>(dlsym(api.handle, "ncclCommInitRank"));
  api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(
      dlsym(api.handle, "ncclAllReduce"));
  api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.handle, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.handle, "ncclGetErrorString"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy;
  return api;
}

constexpr int kNcclFloat32 = 7;  // ncclFloat
constexpr int kNcclSum = 0;      // ncclSum

}  // namespace b200ot

using namespace b200ot;

extern "C" {

int b200ot_sinkhorn_shard_prologue(const float* C, int ldc, int n_local, int m, void* ws, float* s_local, void* stream);
int b200ot_sinkhorn_shard_sweep(const float* C, int ldc, int n_local, int m, int path, void* ws, float* s_local,
                                void* stream);
int b200ot_sinkhorn_shard_finalize(int n_local, int m, void* ws, const float* s_total, int is_prologue, void* stream);

int b200ot_nccl_unique_id(unsigned char* id128_host) {
  if (!id128_host) return B200OT_E_INVALID;
  NcclApi& api = nccl();
  if (!api.ok) return B200OT_E_UNSUPPORTED;
  return api.GetUniqueId(id128_host) == 0 ? 0 : B200OT_E_LAUNCH;
}

int b200ot_nccl_init(const unsigned char* id128_host, int world, int rank, void** comm_out) {
  if (!id128_host || !comm_out || world < 1 || rank < 0 || rank >= world) return B200OT_E_INVALID;
  NcclApi& api = nccl();
  if (!api.ok) return B200OT_E_UNSUPPORTED;
  Id128 id;
  memcpy(id.bytes, id128_host, 128);
  void* comm = nullptr;
  if (api.CommInitRank(&comm, world, id, rank) != 0) return B200OT_E_LAUNCH;
  *comm_out = comm;
  return 0;
}

int b200ot_nccl_destroy(void* comm) {
  if (!comm) return B200OT_E_INVALID;
  NcclApi& api = nccl();
  if (!api.ok) return B200OT_E_UNSUPPORTED;
  return api.CommDestroy(comm) == 0 ? 0 : B200OT_E_LAUNCH;
}

// ---- peer memory over NVLink (CUDA IPC): exchange buffers of the NCCL-free sharded loop ----------------------
int b200ot_peer_alloc(size_t bytes, void** dptr, unsigned char* handle64_host) {
  if (!dptr || !handle64_host || bytes == 0) return B200OT_E_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  B200OT_CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);  // tag 0 never matches
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_last_cuda_error(e, "cudaIpcGetMemHandle");
    (void)cudaGetLastError();
    cudaFree(p);
    return B200OT_E_LAUNCH;
  }
  memcpy(handle64_host, &h, 64);
  *dptr = p;
  return 0;
}

int b200ot_peer_open(const unsigned char* handle64_host, void** dptr) {
  if (!handle64_host || !dptr) return B200OT_E_INVALID;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  void* p = nullptr;
  B200OT_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dptr = p;
  return 0;
}

int b200ot_peer_close(void* dptr) {
  if (!dptr) return B200OT_E_INVALID;
  B200OT_CUDA_OK(cudaIpcCloseMemHandle(dptr));
  return 0;
}

int b200ot_peer_free(void* dptr) {
  if (!dptr) return B200OT_E_INVALID;
  B200OT_CUDA_OK(cudaFree(dptr));
  return 0;
}

// First g update of a row-sharded solve (after b200ot_sinkhorn_setup): column sums, all-reduce, finalize.
int b200ot_sinkhorn_shard_start(const float* C, int ldc, int n_local, int m, void* ws, float* s_buf, void* comm,
                                void* stream) {
  NcclApi& api = nccl();
  if (comm && !api.ok) return B200OT_E_UNSUPPORTED;
  int rc = b200ot_sinkhorn_shard_prologue(C, ldc, n_local, m, ws, s_buf, stream);
  if (rc) return rc;
  if (comm && api.AllReduce(s_buf, s_buf, (size_t)m, kNcclFloat32, kNcclSum, comm, static_cast<cudaStream_t>(stream)) != 0)
    return B200OT_E_LAUNCH;
  return b200ot_sinkhorn_shard_finalize(n_local, m, ws, s_buf, 1, stream);
}

// `iters` iterations: local single-sweep -> fold partials -> ncclAllReduce(m floats, sum) -> finalize, all on
// `stream`.  comm == NULL runs the same loop without the collective (one shard = the whole problem).
int b200ot_sinkhorn_shard_run(const float* C, int ldc, int n_local, int m, int iters, int path, void* ws,
                              float* s_buf, void* comm, void* stream) {
  if (iters < 0) return B200OT_E_INVALID;
  NcclApi& api = nccl();
  if (comm && !api.ok) return B200OT_E_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int i = 0; i < iters; ++i) {
    int rc = b200ot_sinkhorn_shard_sweep(C, ldc, n_local, m, path, ws, s_buf, stream);
    if (rc) return rc;
    if (comm && api.AllReduce(s_buf, s_buf, (size_t)m, kNcclFloat32, kNcclSum, comm, s) != 0) return B200OT_E_LAUNCH;
    rc = b200ot_sinkhorn_shard_finalize(n_local, m, ws, s_buf, 0, stream);
    if (rc) return rc;
  }
  return 0;
}

}  // extern "C"
