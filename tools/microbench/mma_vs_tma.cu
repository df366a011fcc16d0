// Microbenchmark: do the tensor cores' shared-memory operand reads and the TMA's shared-memory writes get in each
// other's way?  One CTA per SM.  An MMA thread issues back-to-back tcgen05.mma (cta_group::1, bf16, M = 128, K = 16,
// N = 64 / 128 / 256; K-major SWIZZLE_64B operands that sit still in shared memory -- the conv kernel's descriptors) and a
// TMA thread streams 8 KB stages of an L2-resident buffer into a ring (the conv kernel's weight loads).  Reported per
// SM: MMA rate alone, TMA rate alone, both together.  Operand bytes read per MMA: 4 KB of A + 32 N bytes of B.
// The A window is either aligned or walks the 25 taps of a 5x5 filter over a strip 100 pixels wide as START-ADDRESS
// SHIFTS (the conv kernel's formulation), in the 64-byte-swizzled [pixel][32 ch] layout or in the unswizzled
// [8-channel group][pixel][8 ch] layout (the conv kernel's a_mode 1).
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a mma_vs_tma.cu -o mma_vs_tma -lcuda
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "../../image-enhance-keras_b200/csrc/ptx.cuh"

using namespace sr;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kStage = 8192, kStages = 8, kGroups = 4, kABytes = 40960;

struct Bars {
  uint64_t full[kStages], empty[kStages], done[kGroups];
  uint32_t tmem_base;
  uint32_t shift_tbl[32];
};

template <int N>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, int buf_stages, int mode, int tma_iters,
                                            int mma_groups, int shift, int layout, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_buf = smem;                       // a strip of 640 pixel rows x 64 B (two K = 16 slices per row)
  uint8_t* b_buf = smem + kABytes;             // N rows x 64 B
  uint8_t* ring = smem + kABytes + 16384;
  Bars* bars = reinterpret_cast<Bars*>(ring + kStages * kStage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (kABytes + 16384) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x < 32) {
    const int tap = threadIdx.x % 25;
    // rows (pixels) into the strip; a row is 64 B in the swizzled layout (>>4: 4 units), 16 B in the unswizzled one (1 unit)
    bars->shift_tbl[threadIdx.x] = shift ? (uint32_t)((tap / 5) * 100 + tap % 5) * (layout ? 1u : 4u) : 0u;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < kGroups; ++i) mbar_init(&bars->done[i], 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 2) {
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  unsigned long long t0 = 0, t1 = 0;
  if (warp == 0 && lane == 0 && (mode & 1)) {          // TMA producer
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int it = 0; it < tma_iters; ++it) {
      const int slot = it % kStages, ph = (it / kStages) & 1;
      mbar_wait(&bars->empty[slot], ph ^ 1);
      mbar_expect_tx(&bars->full[slot], kStage);
      const int st = it % buf_stages;
      tma_load_2d(ring + slot * kStage, &tm, &bars->full[slot], 0, st * 128);
      tma_load_2d(ring + slot * kStage + 4096, &tm, &bars->full[slot], 0, st * 128 + 64);
    }
  } else if (warp == 3 && lane == 0 && (mode & 1)) {   // TMA consumer: frees a stage as soon as it has landed
    for (int it = 0; it < tma_iters; ++it) {
      const int slot = it % kStages, ph = (it / kStages) & 1;
      mbar_wait(&bars->full[slot], ph);
      mbar_arrive(&bars->empty[slot]);
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    out[blockIdx.x * 4 + 1] = t1;
  } else if (warp == 1 && (mode & 2)) {                // MMA issuer: groups of 8 MMAs, at most kGroups groups in flight
    const bool leader = elect_one();
    // swizzled: SBO 512 B (8 rows of 64 B), K = 16 slice = +32 B; unswizzled: SBO 128 B (8 pixels of 16 B), LBO = one
    // 8-channel plane of the strip (640 pixels x 16 B), K = 16 slice = two planes
    const uint32_t kHiA = layout ? ((128u >> 4) | (1u << 14)) : ((512u >> 4) | (1u << 14) | ((uint32_t)SR_LAYOUT_SW64 << 29));
    constexpr uint32_t kHiB = (512u >> 4) | (1u << 14) | ((uint32_t)SR_LAYOUT_SW64 << 29);
    constexpr uint32_t IDESC = umma_idesc(1u, 128u, (uint32_t)N);
    const uint32_t a_lo = (smem_u32(a_buf) >> 4) | (layout ? ((10240u >> 4) << 16) : (1u << 16));
    const uint32_t a_k16 = layout ? (2u * 10240u) >> 4 : 2u;
    const uint32_t b_lo = (smem_u32(b_buf) >> 4) | (1u << 16);
    if (leader) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int t4 = 0;                                   // first of the group's four taps
    for (int g = 0; g < mma_groups; ++g) {
      const int s = g % kGroups, ph = (g / kGroups) & 1;
      if (g >= kGroups) mbar_wait(&bars->done[s], ph ^ 1);
      tc_fence_after();
      if (leader) {
        uint32_t sh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) sh[j] = bars->shift_tbl[t4 + j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint64_t ad = ((uint64_t)kHiA << 32) | (uint64_t)(a_lo + sh[i >> 1] + (i & 1) * a_k16);
          const uint64_t bd = ((uint64_t)kHiB << 32) | (uint64_t)(b_lo + (i & 1) * 2);
          umma_bf16(tmem_base + (uint32_t)((i >> 1) & 1) * N, ad, bd, IDESC, 1u);
        }
        umma_commit(&bars->done[s]);
      }
      t4 = t4 + 4 >= 25 ? 0 : t4 + 4;             // taps 0-3, 4-7, ..., 20-23, (24-27 wrap inside the 32-entry table)
      __syncwarp();
    }
    for (int g = mma_groups; g < mma_groups + kGroups; ++g) {   // drain
      const int s = g % kGroups, ph = (g / kGroups) & 1;
      mbar_wait(&bars->done[s], ph ^ 1);
    }
    if (leader) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      out[blockIdx.x * 4 + 2] = t0;
      out[blockIdx.x * 4 + 3] = t1;
    }
  }
  if (warp == 0 && lane == 0 && (mode & 1)) out[blockIdx.x * 4] = t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int N>
static void run(const CUtensorMap& tm, int buf_stages, int ctas, int shift, int layout, unsigned long long* out_dev) {
  const size_t smem = 1024 + kABytes + 16384 + (size_t)kStages * kStage + sizeof(Bars) + 64;
  CK(cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tma_iters = 6000;                         // 49 MB per CTA: ~0.9 ms alone
  const int mma_groups = 3000 * 128 / N * 2;          // ~1.6 ms of MMAs at the full rate
  unsigned long long* h = (unsigned long long*)malloc(ctas * 4 * sizeof(unsigned long long));
  double res[4][2] = {};
  for (int mode = 1; mode <= 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaMemset(out_dev, 0, ctas * 4 * sizeof(unsigned long long)));
      k<N><<<ctas, 128, smem>>>(tm, buf_stages, mode, tma_iters, mma_groups, shift, layout, out_dev);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, out_dev, ctas * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double tma_ns = 0, mma_ns = 0;
    for (int c = 0; c < ctas; ++c) {
      tma_ns += (double)(h[c * 4 + 1] - h[c * 4]);
      mma_ns += (double)(h[c * 4 + 3] - h[c * 4 + 2]);
    }
    tma_ns /= ctas; mma_ns /= ctas;
    res[mode][0] = (mode & 1) ? (double)tma_iters * kStage / tma_ns : 0.0;                       // GB/s per SM
    res[mode][1] = (mode & 2) ? (double)mma_groups * 8 * 128.0 * N * 16 * 2 / mma_ns / 1e3 : 0.0;  // TFLOP/s per SM
  }
  const double rd = 4096.0 + 32.0 * N;   // operand bytes per MMA
  printf("{\"N\": %d, \"a_layout\": \"%s\", \"a_window\": \"%s\", \"operand_bytes_per_mma\": %d, \"tma_alone_GBps_per_sm\": %.2f, \"mma_alone_TFLOPs_per_sm\": %.3f, "
         "\"mma_alone_operand_GBps_per_sm\": %.1f, \"together_tma_GBps_per_sm\": %.2f, \"together_mma_TFLOPs_per_sm\": %.3f, "
         "\"together_operand_GBps_per_sm\": %.1f, \"chip_mma_alone_TFLOPs\": %.1f, \"chip_mma_together_TFLOPs\": %.1f}\n",
         N, layout ? "unswizzled [8ch][pixel][8ch]" : "swizzle64 [pixel][32ch]", shift ? "5x5 taps as start-address shifts" : "aligned", (int)rd, res[1][0], res[2][1], res[2][1] * 1e3 / (128.0 * N * 16 * 2) * rd, res[3][0], res[3][1],
         res[3][1] * 1e3 / (128.0 * N * 16 * 2) * rd, res[2][1] * ctas, res[3][1] * ctas);
  free(h);
}

int main(int argc, char** argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 148;
  const int buf_kb = 800, buf_stages = buf_kb * 1024 / kStage;
  uint8_t* buf;
  CK(cudaMalloc(&buf, (size_t)buf_stages * kStage));
  CK(cudaMemset(buf, 1, (size_t)buf_stages * kStage));
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  PFN_enc enc = reinterpret_cast<PFN_enc>(p);
  CUtensorMap tm;
  cuuint64_t dims[2] = {32, (cuuint64_t)buf_stages * 128};
  cuuint64_t strides[1] = {64};
  cuuint32_t box[2] = {32, 64}, es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  unsigned long long* out_dev;
  CK(cudaMalloc(&out_dev, ctas * 4 * sizeof(unsigned long long)));
  for (int layout = 0; layout < 2; ++layout)
    for (int shift = 0; shift < 2; ++shift) {
      run<64>(tm, buf_stages, ctas, shift, layout, out_dev);
      run<128>(tm, buf_stages, ctas, shift, layout, out_dev);
      run<256>(tm, buf_stages, ctas, shift, layout, out_dev);
    }
  return 0;
}
