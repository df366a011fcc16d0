// Microbenchmark (companion of tma_rows.cu): does TMA multicast lift the per-SM ingest ceiling?
// Clusters of CS CTAs (2 or 4); every CTA fetches 1/CS of each 8 KB stage with cp.async.bulk ... .multicast::cluster
// into ALL CTAs of the cluster, so every SM RECEIVES 8 KB per stage but REQUESTS only 8/CS KB.  A slot is free when all
// CS consumers have released it (relaxed remote arrives -- in the conv kernel this is tcgen05.commit's multicast).
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tma_multicast.cu -o tma_multicast
// Run:    ./tma_multicast [ctas=148] [stages=9] [buffer_kb=800]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_1d_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* b, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(b)), "r"(rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

constexpr int kStage = 8192;

template <int CS>
__global__ void __launch_bounds__(128, 1) stream_mc_kernel(const uint8_t* buf, int buf_stages, int nstages, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + nstages * kStage);
  uint64_t* empty = full + 32;
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();
  constexpr uint32_t part = kStage / CS;
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
      const int slot = it % nstages, ph = (it / nstages) & 1;
      mbar_wait(&empty[slot], ph ^ 1);
      mbar_expect_tx(&full[slot], kStage);
      const int st = it % buf_stages;
      bulk_1d_mc(smem + slot * kStage + rank * part, buf + (size_t)st * kStage + rank * part, part, &full[slot], (uint16_t)((1u << CS) - 1u));
    }
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + CS) {
    // consumer lane c releases the slot in CTA c of the cluster (one remote arrive per lane, all in flight together)
    const uint32_t c = threadIdx.x - 32;
    for (int it = 0; it < iters; ++it) {
      const int slot = it % nstages, ph = (it / nstages) & 1;
      mbar_wait(&full[slot], ph);
      mbar_arrive_remote_relaxed(&empty[slot], c);
    }
  }
  __syncthreads();
  cluster_sync();
}

template <int CS>
static void run(int ctas, int nstages, int buf_kb, const uint8_t* buf, int buf_stages) {
  const size_t smem = (size_t)nstages * kStage + 1024 + 512;
  CK(cudaFuncSetAttribute(stream_mc_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.attrs = at; cfg.numAttrs = 1;
  int max_clusters = 0;
  cfg.gridDim = dim3(CS);
  CK(cudaOccupancyMaxActiveClusters(&max_clusters, stream_mc_kernel<CS>, &cfg));
  int n = ctas / CS * CS;
  if (n > max_clusters * CS) n = max_clusters * CS;
  cfg.gridDim = dim3(n);
  const int iters = 4000;
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, stream_mc_kernel<CS>, buf, buf_stages, nstages, iters));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  const double bytes = (double)n * iters * kStage;
  printf("{\"mode\": \"bulk1d multicast over %d CTAs\", \"max_resident_clusters\": %d, \"ctas\": %d, \"stages\": %d, \"buffer_kb\": %d, \"ms\": %.4f, "
         "\"received_GBps_per_sm\": %.2f, \"requested_GBps_per_sm\": %.2f, \"chip_received_TBps\": %.3f}\n",
         CS, max_clusters, n, nstages, buf_kb, ms, bytes / ms / 1e6 / n, bytes / ms / 1e6 / n / CS, bytes / ms / 1e9);
}

int main(int argc, char** argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 148;
  const int nstages = argc > 2 ? atoi(argv[2]) : 9;
  const int buf_kb = argc > 3 ? atoi(argv[3]) : 800;
  const int buf_stages = buf_kb * 1024 / kStage;
  uint8_t* buf;
  CK(cudaMalloc(&buf, (size_t)buf_stages * kStage));
  CK(cudaMemset(buf, 1, (size_t)buf_stages * kStage));
  run<2>(ctas, nstages, buf_kb, buf, buf_stages);
  run<4>(ctas, nstages, buf_kb, buf, buf_stages);
  return 0;
}
