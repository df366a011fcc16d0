// The one exchange step of the path: data-parallel training averages the gradients of the minibatch
// shards and applies compile(Adam(1e-4, 0.9)) (models.py:1212-1213) to identical replicas.  Instead of
// ncclAllReduce -> Adam as two passes, ONE kernel per rank does reduce-scatter + Adam + all-gather over NVLink peer
// memory (one process per GPU, buffers shared with CUDA IPC):
//
//   rank r owns the contiguous shard [lo_r, hi_r) of the flat fp32 arena (21.8 M floats / world);
//   ready flags  : every rank tells every peer "my gradient arena of step e is complete";
//   reduce       : g[i] = sum over ranks (in rank order 0..W-1: deterministic) of grads_rank[i], i in my shard,
//                  read straight from the peers' arenas (W-1 remote reads of 1/W of the arena each);
//   Adam         : Keras-2 update of p, m, v on the shard only (the optimizer state is sharded: 1/W of the work);
//   all-gather   : the new p[i] is stored into EVERY rank's parameter arena (W-1 remote writes);
//   done flags   : the last block of every rank tells the peers "my shard is written everywhere" and waits for theirs,
//                  so when the kernel retires the whole local parameter arena is current and the weight repack
//                  that follows in stream order reads it.
//
// Per rank 2 * (W-1)/W * 87.4 MB cross the links (both directions at once) instead of the ring all-reduce's
// 2 * (W-1)/W * 87.4 MB plus a separate 611 MB Adam pass over the whole arena on every rank.
// Waits are bounded (globaltimer): a missing peer sets a status bit instead of hanging the GPU.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>

#include "internal.h"

namespace sr {
namespace {

constexpr int kMaxPeers = SR_EXCHANGE_MAX_RANKS;
// signal pad (uint32 words): [0,8) ready[src], [8,16) done[src], 16 block counter, 17 epoch, 18 status
constexpr int kReady = 0, kDone = 8, kCounter = 16, kEpoch = 17, kStatus = 18;
constexpr int kThreads = 256;

struct ExParams {
  float* grads[kMaxPeers];
  float* params[kMaxPeers];
  unsigned* sig[kMaxPeers];
  int rank, world;
  size_t n;            // floats in the arena
  size_t lo4, hi4;     // this rank's shard in float4 units
  float* m;
  float* v;
  float lr_t, b1, b2, eps, gscale;
  unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer arenas change every step and are written by other GPUs: never through the non-coherent path
__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_sys_f1(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// false on time-out (status bit set)
__device__ __forceinline__ bool wait_flag(const unsigned* flag, unsigned epoch, unsigned* status,
                                          unsigned long long timeout_ns) {
  const unsigned long long t0 = now_ns();
  unsigned spins = 0;
  while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
    if ((++spins & 0x3FFu) == 0 && now_ns() - t0 > timeout_ns) {
      atomicOr(status, 1u);
      return false;
    }
    __nanosleep(64);
  }
  return true;
}

// same expression as adam_kernel (elementwise.cu): a world-1 exchange equals sr_adam_step bit for bit
__device__ __forceinline__ float adam_one(float g, float& m, float& v, float p, const ExParams& P) {
  const float gi = g * P.gscale;
  const float mi = P.b1 * m + (1.f - P.b1) * gi;
  const float vi = P.b2 * v + (1.f - P.b2) * gi * gi;
  m = mi;
  v = vi;
  return p - P.lr_t * mi / (sqrtf(vi) + P.eps);
}

// W = world size, U = float4 elements per thread and iteration: U * W remote 16-byte loads are in flight per thread
// (NVLink latency is paid once per iteration), 8 for every world size.
template <int W, int U>
__global__ void __launch_bounds__(kThreads) exchange_adam_kernel(const ExParams P) {
  unsigned* my = P.sig[P.rank];
  __shared__ unsigned s_epoch;
  __shared__ int s_last;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(my + kEpoch) + 1u;
  __syncthreads();
  const unsigned e = s_epoch;
  // ready: the gradient arena of this rank was completed by earlier kernels of this stream
  if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(P.sig[threadIdx.x] + kReady + P.rank, e);
  if (threadIdx.x < W) wait_flag(my + kReady + threadIdx.x, e, my + kStatus, P.timeout_ns);
  __syncthreads();

  const size_t stride = (size_t)gridDim.x * blockDim.x * U;
  for (size_t base = P.lo4 + (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < P.hi4; base += stride) {
    float4 q[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i4 = base + (size_t)u * kThreads;
      if (i4 < P.hi4) {
#pragma unroll
        for (int r = 0; r < W; ++r) q[u][r] = ld_sys_f4(P.grads[r] + i4 * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i4 = base + (size_t)u * kThreads;
      if (i4 >= P.hi4) break;
      const size_t i = i4 * 4;
      float4 g = q[u][0];
#pragma unroll
      for (int r = 1; r < W; ++r) g.x += q[u][r].x, g.y += q[u][r].y, g.z += q[u][r].z, g.w += q[u][r].w;
      float4 m = *reinterpret_cast<const float4*>(P.m + i);
      float4 v = *reinterpret_cast<const float4*>(P.v + i);
      float4 p = *reinterpret_cast<const float4*>(P.params[P.rank] + i);
      p.x = adam_one(g.x, m.x, v.x, p.x, P);
      p.y = adam_one(g.y, m.y, v.y, p.y, P);
      p.z = adam_one(g.z, m.z, v.z, p.z, P);
      p.w = adam_one(g.w, m.w, v.w, p.w, P);
      *reinterpret_cast<float4*>(P.m + i) = m;
      *reinterpret_cast<float4*>(P.v + i) = v;
#pragma unroll
      for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(P.params[r] + i) = p;
    }
  }
  // the last n % 4 floats belong to the last rank
  if (P.rank == W - 1 && blockIdx.x == 0 && threadIdx.x < (int)(P.n & 3)) {
    const size_t i = (P.n & ~(size_t)3) + threadIdx.x;
    float g = ld_sys_f1(P.grads[0] + i);
    for (int r = 1; r < W; ++r) g += ld_sys_f1(P.grads[r] + i);
    float m = P.m[i], v = P.v[i];
    const float p = adam_one(g, m, v, P.params[P.rank][i], P);
    P.m[i] = m;
    P.v[i] = v;
    for (int r = 0; r < W; ++r) P.params[r][i] = p;
  }
  // done: every block's remote stores are performed before the last block raises the flag
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(my + kCounter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x < W) {
    st_release_sys(P.sig[threadIdx.x] + kDone + P.rank, e);
    wait_flag(my + kDone + threadIdx.x, e, my + kStatus, P.timeout_ns);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    my[kCounter] = 0;
    *reinterpret_cast<volatile unsigned*>(my + kEpoch) = e;
    __threadfence();
  }
}

// Barrier over the ranks' signal pads: "everything this rank launched before on this stream (its stores into peer
// memory included) is done" -- one warp signals every peer and waits for every peer.
struct BarrierParams {
  unsigned* sig[kMaxPeers];
  int rank, world;
  unsigned long long timeout_ns;
};

__global__ void peer_barrier_kernel(const BarrierParams P) {
  unsigned* my = P.sig[P.rank];
  const unsigned e = *reinterpret_cast<volatile unsigned*>(my + kEpoch) + 1u;
  __threadfence_system();
  if (threadIdx.x < P.world) {
    st_release_sys(P.sig[threadIdx.x] + kReady + P.rank, e);
    wait_flag(my + kReady + threadIdx.x, e, my + kStatus, P.timeout_ns);
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned*>(my + kEpoch) = e;
    __threadfence();
  }
}

}  // namespace
}  // namespace sr

using namespace sr;

struct sr_exchange {
  ExParams P;
  double timeout_ms;
};

extern "C" size_t sr_exchange_signal_bytes(void) { return 32 * sizeof(unsigned); }

extern "C" int sr_ipc_export(const void* dev_ptr, unsigned char* handle, size_t* offset) {
  if (!dev_ptr || !handle || !offset) return set_error(SR_ERR_INVALID, "sr_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == SR_IPC_HANDLE_BYTES, "handle size");
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, dev_ptr);
  if (e != cudaSuccess) return set_cuda_error(e, "sr_ipc_export: cudaPointerGetAttributes");
  if (attr.type != cudaMemoryTypeDevice) return set_error(SR_ERR_INVALID, "sr_ipc_export: not a device pointer");
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr));   // the handle names the whole cudaMalloc allocation
  if (e != cudaSuccess) return set_cuda_error(e, "sr_ipc_export: cudaIpcGetMemHandle");
  // offset of dev_ptr inside that allocation: cuMemGetAddressRange, fetched through the runtime (no libcuda link)
  typedef CUresult (*PFN_range)(CUdeviceptr*, size_t*, CUdeviceptr);
  static PFN_range range_fn = nullptr;
  if (!range_fn) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return set_error(SR_ERR_CUDA, "sr_ipc_export: cuMemGetAddressRange is not available");
    range_fn = reinterpret_cast<PFN_range>(fp);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  if (range_fn(&base, &size, reinterpret_cast<CUdeviceptr>(dev_ptr)) != CUDA_SUCCESS)
    return set_error(SR_ERR_CUDA, "sr_ipc_export: cuMemGetAddressRange failed");
  memcpy(handle, &h, sizeof h);
  *offset = (size_t)(reinterpret_cast<CUdeviceptr>(dev_ptr) - base);
  return SR_OK;
}

extern "C" int sr_ipc_open(const unsigned char* handle, size_t offset, void** dev_ptr) {
  if (!handle || !dev_ptr) return set_error(SR_ERR_INVALID, "sr_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  void* base = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return set_cuda_error(e, "sr_ipc_open: cudaIpcOpenMemHandle");
  *dev_ptr = reinterpret_cast<char*>(base) + offset;
  return SR_OK;
}

extern "C" int sr_ipc_close(void* dev_ptr, size_t offset) {
  if (!dev_ptr) return SR_OK;
  cudaError_t e = cudaIpcCloseMemHandle(reinterpret_cast<char*>(dev_ptr) - offset);
  if (e != cudaSuccess) return set_cuda_error(e, "sr_ipc_close");
  return SR_OK;
}

extern "C" int sr_exchange_shard(size_t n, int rank, int world, size_t* lo, size_t* hi) {
  if (world < 1 || rank < 0 || rank >= world || !lo || !hi) return set_error(SR_ERR_INVALID, "sr_exchange_shard: bad rank");
  const size_t n4 = n / 4;
  *lo = 4 * (n4 * (size_t)rank / world);
  *hi = rank == world - 1 ? n : 4 * (n4 * (size_t)(rank + 1) / world);
  return SR_OK;
}

extern "C" int sr_exchange_create(int rank, int world, size_t n, float* const* grads, float* const* params,
                                  void* const* signals, sr_exchange** out) {
  if (!grads || !params || !signals || !out) return set_error(SR_ERR_INVALID, "sr_exchange_create: null argument");
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return set_error(SR_ERR_INVALID, "sr_exchange_create: world must be 1..8 and 0 <= rank < world");
  for (int r = 0; r < world; ++r) {
    if (!grads[r] || !params[r] || !signals[r]) return set_error(SR_ERR_INVALID, "sr_exchange_create: null peer pointer");
    if ((reinterpret_cast<uintptr_t>(grads[r]) | reinterpret_cast<uintptr_t>(params[r])) & 15)
      return set_error(SR_ERR_INVALID, "sr_exchange_create: arenas must be 16-byte aligned");
  }
  sr_exchange* ex = new (std::nothrow) sr_exchange();
  if (!ex) return set_error(SR_ERR_NOMEM, "sr_exchange_create: out of memory");
  memset(&ex->P, 0, sizeof ex->P);
  for (int r = 0; r < world; ++r) {
    ex->P.grads[r] = grads[r];
    ex->P.params[r] = params[r];
    ex->P.sig[r] = reinterpret_cast<unsigned*>(signals[r]);
  }
  ex->P.rank = rank, ex->P.world = world, ex->P.n = n;
  size_t lo, hi;
  sr_exchange_shard(n, rank, world, &lo, &hi);
  ex->P.lo4 = lo / 4;
  ex->P.hi4 = (rank == world - 1 ? (n & ~(size_t)3) : hi) / 4;
  ex->timeout_ms = 30000.0;
  *out = ex;
  return SR_OK;
}

extern "C" void sr_exchange_destroy(sr_exchange* ex) { delete ex; }

extern "C" int sr_exchange_set_timeout_ms(sr_exchange* ex, double ms) {
  if (!ex || !(ms > 0)) return set_error(SR_ERR_INVALID, "sr_exchange_set_timeout_ms: bad argument");
  ex->timeout_ms = ms;
  return SR_OK;
}

extern "C" int sr_exchange_adam_step(sr_exchange* ex, float* m, float* v, int t, float lr, float beta1, float beta2,
                                     float eps, float grad_scale, int max_blocks, void* stream) {
  if (!ex || !m || !v) return set_error(SR_ERR_INVALID, "sr_exchange_adam_step: null argument");
  if (t < 1) return set_error(SR_ERR_INVALID, "sr_exchange_adam_step: t must be >= 1");
  ExParams P = ex->P;
  P.m = m, P.v = v;
  P.lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  P.b1 = beta1, P.b2 = beta2, P.eps = eps, P.gscale = grad_scale;
  P.timeout_ns = (unsigned long long)(ex->timeout_ms * 1e6);
  // every block stays resident while it waits for the peers' flags: 2 blocks of 256 threads per SM (80 registers)
  int grid = max_blocks > 0 ? max_blocks : 148 * 2;
  if (grid > 148 * 2) grid = 148 * 2;
  const size_t work = P.hi4 > P.lo4 ? P.hi4 - P.lo4 : 1;
  const size_t need = (work + kThreads - 1) / kThreads;
  if ((size_t)grid > need) grid = (int)need;
  cudaStream_t st = as_stream(stream);
  switch (P.world) {
    case 1: exchange_adam_kernel<1, 4><<<grid, kThreads, 0, st>>>(P); break;
    case 2: exchange_adam_kernel<2, 4><<<grid, kThreads, 0, st>>>(P); break;
    case 3: exchange_adam_kernel<3, 2><<<grid, kThreads, 0, st>>>(P); break;
    case 4: exchange_adam_kernel<4, 2><<<grid, kThreads, 0, st>>>(P); break;
    case 5: exchange_adam_kernel<5, 1><<<grid, kThreads, 0, st>>>(P); break;
    case 6: exchange_adam_kernel<6, 1><<<grid, kThreads, 0, st>>>(P); break;
    case 7: exchange_adam_kernel<7, 1><<<grid, kThreads, 0, st>>>(P); break;
    default: exchange_adam_kernel<8, 1><<<grid, kThreads, 0, st>>>(P); break;
  }
  return check_launch("exchange_adam_kernel");
}

extern "C" int sr_exchange_status(sr_exchange* ex, void* stream, int* timed_out) {
  if (!ex || !timed_out) return set_error(SR_ERR_INVALID, "sr_exchange_status: null argument");
  unsigned s = 0;
  cudaError_t e = cudaMemcpyAsync(&s, ex->P.sig[ex->P.rank] + kStatus, sizeof s, cudaMemcpyDeviceToHost, as_stream(stream));
  if (e == cudaSuccess) e = cudaStreamSynchronize(as_stream(stream));
  if (e != cudaSuccess) return set_cuda_error(e, "sr_exchange_status");
  *timed_out = (int)(s & 1u);
  return SR_OK;
}

struct sr_peer_barrier {
  BarrierParams P;
  double timeout_ms;
};

extern "C" int sr_peer_barrier_create(int rank, int world, void* const* signals, sr_peer_barrier** out) {
  if (!signals || !out) return set_error(SR_ERR_INVALID, "sr_peer_barrier_create: null argument");
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return set_error(SR_ERR_INVALID, "sr_peer_barrier_create: world must be 1..8 and 0 <= rank < world");
  sr_peer_barrier* b = new (std::nothrow) sr_peer_barrier();
  if (!b) return set_error(SR_ERR_NOMEM, "sr_peer_barrier_create: out of memory");
  memset(&b->P, 0, sizeof b->P);
  for (int r = 0; r < world; ++r) {
    if (!signals[r]) {
      delete b;
      return set_error(SR_ERR_INVALID, "sr_peer_barrier_create: null peer pointer");
    }
    b->P.sig[r] = reinterpret_cast<unsigned*>(signals[r]);
  }
  b->P.rank = rank, b->P.world = world;
  b->timeout_ms = 30000.0;
  *out = b;
  return SR_OK;
}

extern "C" void sr_peer_barrier_destroy(sr_peer_barrier* b) { delete b; }

extern "C" int sr_peer_barrier_set_timeout_ms(sr_peer_barrier* b, double ms) {
  if (!b || !(ms > 0)) return set_error(SR_ERR_INVALID, "sr_peer_barrier_set_timeout_ms: bad argument");
  b->timeout_ms = ms;
  return SR_OK;
}

extern "C" int sr_peer_barrier_arrive_wait(sr_peer_barrier* b, void* stream) {
  if (!b) return set_error(SR_ERR_INVALID, "sr_peer_barrier_arrive_wait: null argument");
  BarrierParams P = b->P;
  P.timeout_ns = (unsigned long long)(b->timeout_ms * 1e6);
  peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(P);
  return check_launch("peer_barrier_kernel");
}

extern "C" int sr_peer_barrier_status(sr_peer_barrier* b, void* stream, int* timed_out) {
  if (!b || !timed_out) return set_error(SR_ERR_INVALID, "sr_peer_barrier_status: null argument");
  unsigned s = 0;
  cudaError_t e = cudaMemcpyAsync(&s, b->P.sig[b->P.rank] + kStatus, sizeof s, cudaMemcpyDeviceToHost, as_stream(stream));
  if (e == cudaSuccess) e = cudaStreamSynchronize(as_stream(stream));
  if (e != cudaSuccess) return set_cuda_error(e, "sr_peer_barrier_status");
  *timed_out = (int)(s & 1u);
  return SR_OK;
}
