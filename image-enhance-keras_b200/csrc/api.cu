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

namespace sr {
static int g_pdl = 1;
bool pdl_enabled() { return g_pdl != 0; }
}  // namespace sr

// 1 (default): the tensor-core kernels are launched with programmatic stream serialization (see internal.h);
// 0: plain stream-ordered launches.  Same results either way; takes effect for launches (and graph captures) made
// afterwards.  Returns the previous setting.
extern "C" int sr_set_pdl(int enabled) {
  const int prev = sr::g_pdl;
  sr::g_pdl = enabled ? 1 : 0;
  return prev;
}

extern "C" int sr_version(void) { return 100; }

// 1 when the library was compiled with -DSR_DEV_SWITCHES (timing-only environment switches that can change
// results); bench.py and the tests refuse such a build.
extern "C" int sr_dev_switches(void) {
#ifdef SR_DEV_SWITCHES
  return 1;
#else
  return 0;
#endif
}

#ifdef SR_DEV_SWITCHES
namespace sr {
static unsigned long long* g_timeline = nullptr;
unsigned long long* dev_timeline() { return g_timeline; }
}  // namespace sr
#endif

// Development build only: device buffer (16 x 8 bytes per CTA) that conv plans created afterwards stamp their phase
// clocks into (tools/probe_timeline.py); NULL turns it off.  The release library refuses.
extern "C" int sr_dev_set_timeline(void* buf) {
#ifdef SR_DEV_SWITCHES
  sr::g_timeline = reinterpret_cast<unsigned long long*>(buf);
  return SR_OK;
#else
  (void)buf;
  return sr::set_error(SR_ERR_UNSUPPORTED, "sr_dev_set_timeline: not a development build (-DSR_DEV_SWITCHES)");
#endif
}

extern "C" int sr_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

// Layout check for bindings: sizeof of the ABI structs as this library was compiled (a binding whose struct
// declaration fell behind the header would otherwise hand the library a short struct).
extern "C" size_t sr_abi_struct_size(int which) {
  switch (which) {
    case 0: return sizeof(sr_conv_desc);
    case 1: return sizeof(sr_conv_plan_info_t);
    case 2: return sizeof(sr_pack_item);
    case 3: return sizeof(sr_wgrad_desc);
    case 4: return sizeof(sr_wgrad_plan_info_t);
    case 5: return sizeof(sr_score_result);
    case 6: return sizeof(sr_model_config);
    case 7: return sizeof(sr_forward_desc);
    case 8: return sizeof(sr_train_desc);
    case 9: return sizeof(sr_model_run_info);
    case 10: return sizeof(sr_stitch_tile);
    case 11: return sizeof(sr_score_item);
    default: return 0;
  }
}
