// Internal helpers shared by the sr100 translation units (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "../../include/sr100.h"

namespace sr {

int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);

inline int check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, where);
  return SR_OK;
}

// Development switches (timing experiments that make kernels skip operand loads or MMAs, ring / cost overrides):
// compiled in only with -DSR_DEV_SWITCHES (make DEV=1).  The release library reads no environment variable that
// can change a result: SR_DBG() is the constant false and dev_getenv() the constant null.
#ifdef SR_DEV_SWITCHES
#define SR_DBG(P, bit) (((P).dbg & (bit)) != 0)
inline const char* dev_getenv(const char* name) { return getenv(name); }
// phase stamps: P.timeline[cta * 16 + idx] = SM clock (idx 14 / 15: globaltimer at entry / exit)
#define SR_STAMP(P, idx)                                                                      \
  do {                                                                                        \
    if ((P).timeline) (P).timeline[(size_t)blockIdx.x * 16 + (idx)] = (unsigned long long)clock64(); \
  } while (0)
#define SR_STAMP_NS(P, idx)                                                                   \
  do {                                                                                        \
    if ((P).timeline) {                                                                       \
      unsigned long long t_;                                                                  \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                  \
      (P).timeline[(size_t)blockIdx.x * 16 + (idx)] = t_;                                     \
    }                                                                                         \
  } while (0)
unsigned long long* dev_timeline();
#else
#define SR_DBG(P, bit) false
#define SR_STAMP(P, idx) do { } while (0)
#define SR_STAMP_NS(P, idx) do { } while (0)
inline const char* dev_getenv(const char*) { return nullptr; }
inline unsigned long long* dev_timeline() { return nullptr; }
#endif

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch of the tensor-core kernels (sr_set_pdl; default on): the kernel may become resident
// while its predecessor in the stream drains, so its set-up (barriers, TMEM allocation, descriptor prefetch) and the
// launch latency leave the critical path; the kernels call griddep_wait() before their first global access.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: `done` is the caller's per-kernel bit
// mask of the devices it has been applied on (a process normally drives one GPU, but nothing here assumes it).
template <typename Kern>
inline int ensure_dynamic_smem(Kern kern, int bytes, unsigned long long* done, const char* what) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (*done & bit) return SR_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return set_cuda_error(e, what);
  *done |= bit;
  return SR_OK;
}

inline unsigned int grid_for(size_t work_items, int block, int max_blocks = 148 * 16) {
  size_t g = (work_items + block - 1) / block;
  if (g < 1) g = 1;
  if (g > (size_t)max_blocks) g = max_blocks;
  return (unsigned int)g;
}

}  // namespace sr
