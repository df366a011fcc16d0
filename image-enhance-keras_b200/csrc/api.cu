// Error reporting and library identification for the sr100 C ABI (include/sr100.h).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "internal.h"

namespace sr {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg ? msg : "");
  return code;
}

int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof g_err, "%s: %s", where ? where : "cuda", cudaGetErrorString(e));
  return SR_ERR_CUDA;
}

}  // namespace sr

extern "C" const char* sr_last_error_string(void) { return sr::g_err; }

extern "C" int sr_version(void) { return 100; }

extern "C" int sr_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
