// Microbenchmark: how fast can every SM stream the SAME small L2-resident buffer (the conv kernel's weight stages)
// into shared memory, as a function of how the copy is expressed?
//   mode 0  cp.async.bulk.tensor 2D, 64-byte rows (box 64 rows x 64 B, SWIZZLE_64B)  -- what conv_tc.cu does today
//   mode 1  cp.async.bulk.tensor 2D, 128-byte rows (box 32 rows x 128 B, SWIZZLE_128B)
//   mode 2  cp.async.bulk (1D, no tensor map), 4 KB per copy
//   mode 3  cp.async.bulk (1D), 8 KB per copy
// One producer thread per CTA keeps a ring of 8 KB stages full; a consumer thread frees a stage as soon as it has
// landed.  Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tma_rows.cu -o tma_rows
// Run:    ./tma_rows [ctas=148] [stages=9] [buffer_kb=800] [private=0]
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
__device__ __forceinline__ void bulk_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kStage = 8192;

__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm128,
                                                        const uint8_t* buf, int buf_stages, int priv, int mode, int nstages, int iters,
                                                        unsigned long long* t_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + nstages * kStage);
  uint64_t* empty = full + 32;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int base_stage = priv ? (blockIdx.x * 7) % buf_stages : 0;   // private: every CTA starts elsewhere in a big buffer
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int it = 0; it < iters; ++it) {
      const int slot = it % nstages, ph = (it / nstages) & 1;
      mbar_wait(&empty[slot], ph ^ 1);
      mbar_expect_tx(&full[slot], kStage);
      const int st = (base_stage + it) % buf_stages;
      uint8_t* dst = smem + slot * kStage;
      if (mode == 0) {
        tma_2d(dst, &tm64, &full[slot], 0, st * 128);
        tma_2d(dst + 4096, &tm64, &full[slot], 0, st * 128 + 64);
      } else if (mode == 1) {
        tma_2d(dst, &tm128, &full[slot], 0, st * 64);
        tma_2d(dst + 4096, &tm128, &full[slot], 0, st * 64 + 32);
      } else if (mode == 2) {
        bulk_1d(dst, buf + (size_t)st * kStage, 4096, &full[slot]);
        bulk_1d(dst + 4096, buf + (size_t)st * kStage + 4096, 4096, &full[slot]);
      } else {
        bulk_1d(dst, buf + (size_t)st * kStage, 8192, &full[slot]);
      }
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < iters; ++it) {
      const int slot = it % nstages, ph = (it / nstages) & 1;
      mbar_wait(&full[slot], ph);
      mbar_arrive(&empty[slot]);
    }
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    t_out[blockIdx.x * 2 + 1] = t1;
  }
  if (threadIdx.x == 0) t_out[blockIdx.x * 2] = t0;
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 148;
  const int nstages = argc > 2 ? atoi(argv[2]) : 9;
  const int buf_kb = argc > 3 ? atoi(argv[3]) : 800;
  const int priv = argc > 4 ? atoi(argv[4]) : 0;
  const int buf_stages = buf_kb * 1024 / kStage;
  uint8_t* buf;
  CK(cudaMalloc(&buf, (size_t)buf_stages * kStage));
  CK(cudaMemset(buf, 1, (size_t)buf_stages * kStage));
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  PFN_enc enc = reinterpret_cast<PFN_enc>(p);
  CUtensorMap tm64, tm128;
  {
    cuuint64_t dims[2] = {32, (cuuint64_t)buf_stages * 128};
    cuuint64_t strides[1] = {64};
    cuuint32_t box[2] = {32, 64}, es[2] = {1, 1};
    CUresult r = enc(&tm64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode 64 failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)buf_stages * 64};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, 32}, es[2] = {1, 1};
    CUresult r = enc(&tm128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode 128 failed %d\n", (int)r); return 1; }
  }
  unsigned long long* t_out;
  CK(cudaMalloc(&t_out, ctas * 2 * sizeof(unsigned long long)));
  const size_t smem = (size_t)nstages * kStage + 1024 + 512;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const int iters = 4000;   // 32 MB per CTA
  const char* names[4] = {"tensor2d_64B_rows_sw64", "tensor2d_128B_rows_sw128", "bulk1d_4KB", "bulk1d_8KB"};
  for (int rep = 0; rep < 2; ++rep)
    for (int mode = 0; mode < 4; ++mode) {
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0));
      stream_kernel<<<ctas, 128, smem>>>(tm64, tm128, buf, buf_stages, priv, mode, nstages, iters, t_out);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double bytes = (double)ctas * iters * kStage;
      if (rep == 1)
        printf("{\"mode\": \"%s\", \"ctas\": %d, \"stages\": %d, \"buffer_kb\": %d, \"private\": %d, \"ms\": %.4f, \"chip_TBps\": %.3f, \"GBps_per_sm\": %.2f}\n",
               names[mode], ctas, nstages, buf_kb, priv, ms, bytes / ms / 1e9, bytes / ms / 1e6 / ctas);
    }
  return 0;
}
