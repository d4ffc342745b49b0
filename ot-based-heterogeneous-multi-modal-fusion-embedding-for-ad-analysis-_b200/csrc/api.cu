// Library-wide C ABI helpers: version, status strings, last CUDA error.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace b200ot {

static thread_local char g_last_error[512] = "";

void set_last_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where ? where : "?",
           cudaGetErrorString(e), cudaGetErrorName(e));
}

int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  return dev < 0 ? 0 : dev % kMaxDeviceSlots;
}

int sm_count() {
  static int cached[kMaxDeviceSlots] = {};
  const int slot = device_slot();
  if (cached[slot]) return cached[slot];
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached[slot] = n;
  return n;
}

}  // namespace b200ot

extern "C" {

int b200ot_version(void) { return B200OT_VERSION; }

const char* b200ot_strerror(int code) {
  switch (code) {
    case B200OT_OK: return "ok";
    case B200OT_E_INVALID: return "invalid argument";
    case B200OT_E_WORKSPACE: return "workspace too small";
    case B200OT_E_LAUNCH: return "CUDA launch or runtime failure";
    case B200OT_E_UNSUPPORTED: return "unsupported shape or configuration";
    case B200OT_E_NUMERIC: return "non-finite or vanished sum on the fast path";
    default: return "unknown b200ot status";
  }
}

const char* b200ot_last_cuda_error(void) { return b200ot::g_last_error; }

}  // extern "C"
