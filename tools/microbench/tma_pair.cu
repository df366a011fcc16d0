// Microbenchmark (companion of tma_rows.cu): the weight loads of the CTA-pair conv kernel -- 2D tensor-map boxes of
// 64 rows x 64 B issued by BOTH CTAs of a two-CTA cluster with .cta_group::2, their completion bytes credited to ONE
// mbarrier in the leader CTA (rank 0) -- against the same loads credited to a barrier in the issuing CTA.
//   mode 0  plain loads, own barrier, independent CTAs inside the cluster          (reference: tma_rows mode 0)
//   mode 1  .cta_group::2 loads, both CTAs credit rank 0's barrier; rank 0's consumer frees the slot in both CTAs
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tma_pair.cu -o tma_pair -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const void* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_2d_pair(void* dst, const void* tm, uint32_t bar_cluster, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void arrive_remote_relaxed(uint32_t a) { asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory"); }

constexpr int kStage = 8192;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const __grid_constant__ CUtensorMap tm64, int buf_stages, int mode, int nstages, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + nstages * kStage);
  uint64_t* empty = full + 32;
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();
  if (threadIdx.x == 0) {          // producer (both CTAs)
    for (int it = 0; it < iters; ++it) {
      const int slot = it % nstages, ph = (it / nstages) & 1;
      mbar_wait(&empty[slot], ph ^ 1);
      const int st = it % buf_stages;
      uint8_t* dst = smem + slot * kStage;
      if (mode == 0) {
        mbar_expect_tx(&full[slot], kStage);
        tma_2d(dst, &tm64, &full[slot], 0, st * 128);
        tma_2d(dst + 4096, &tm64, &full[slot], 0, st * 128 + 64);
      } else {
        if (rank == 0) mbar_expect_tx(&full[slot], 2 * kStage);     // the leader's barrier collects both CTAs' bytes
        const uint32_t bar = mapa(smem_u32(&full[slot]), 0);
        tma_2d_pair(dst, &tm64, bar, 0, st * 128);
        tma_2d_pair(dst + 4096, &tm64, bar, 0, st * 128 + 64);
      }
    }
  } else if (threadIdx.x == 32 || threadIdx.x == 33) {   // consumer
    if (mode == 0) {
      if (threadIdx.x == 32)
        for (int it = 0; it < iters; ++it) {
          const int slot = it % nstages, ph = (it / nstages) & 1;
          mbar_wait(&full[slot], ph);
          mbar_arrive(&empty[slot]);
        }
    } else if (rank == 0) {          // leader only: lane c frees the slot in CTA c (tcgen05.commit's multicast in the conv kernel)
      const uint32_t c = threadIdx.x - 32;
      for (int it = 0; it < iters; ++it) {
        const int slot = it % nstages, ph = (it / nstages) & 1;
        mbar_wait(&full[slot], ph);
        arrive_remote_relaxed(mapa(smem_u32(&empty[slot]), c));
      }
    }
  }
  __syncthreads();
  cluster_sync();
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 148;
  const int nstages = argc > 2 ? atoi(argv[2]) : 9;
  const int buf_kb = argc > 3 ? atoi(argv[3]) : 800;
  const int buf_stages = buf_kb * 1024 / kStage;
  uint8_t* buf;
  CK(cudaMalloc(&buf, (size_t)buf_stages * kStage));
  CK(cudaMemset(buf, 1, (size_t)buf_stages * kStage));
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  PFN_enc enc = reinterpret_cast<PFN_enc>(p);
  CUtensorMap tm64;
  cuuint64_t dims[2] = {32, (cuuint64_t)buf_stages * 128};
  cuuint64_t strides[1] = {64};
  cuuint32_t box[2] = {32, 64}, es[2] = {1, 1};
  CUresult r = enc(&tm64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const size_t smem = (size_t)nstages * kStage + 1024 + 512;
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 4000;
  const char* names[2] = {"plain loads, own barrier", "cta_group::2 loads, leader's barrier"};
  for (int rep = 0; rep < 2; ++rep)
    for (int mode = 0; mode < 2; ++mode) {
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0));
      pair_kernel<<<ctas & ~1, 128, smem>>>(tm64, buf_stages, mode, nstages, iters);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double bytes = (double)(ctas & ~1) * iters * kStage;
      if (rep == 1)
        printf("{\"mode\": \"%s\", \"ctas\": %d, \"stages\": %d, \"ms\": %.4f, \"chip_TBps\": %.3f, \"GBps_per_sm\": %.2f}\n",
               names[mode], ctas & ~1, nstages, ms, bytes / ms / 1e9, bytes / ms / 1e6 / (ctas & ~1));
    }
  return 0;
}
