// tcgen05 filter-gradient (wgrad) kernel for the 128->128 convolutions of the DifvdsrDouble stack.
// Reference semantics: the gradient Keras/TensorFlow computes for Conv2D(padding='same') kernels during
// fit (models.py:131-157 with compile(mse, Adam) models.py:1212-1213; layers models.py:1231-1270):
//     dW[ky][kx][ci][co] = scale * sum_{n,y,x} X[n, y+ky-p, x+kx-p, ci] * G[n, y, x, co]      (zero outside)
//
// Formulation: per filter tap a GEMM  D_tap[ci][co] = sum_pixels X_shift[pixel][ci] * G[pixel][co]
// with M = ci = 128, N = co = 128, K = pixels.  Both operands are the NHWC tensors themselves, i.e. the
// reduction index (pixel) is the SLOW index in memory: they are fed to the tensor cores MN-major
// (UMMA descriptor a_major = b_major = 1) from TMA 128B-swizzled boxes of [pixels][64 channels].
//   * a CTA owns a tap group (<= 4 taps = 4 x 128 TMEM columns, fp32) for its whole lifetime and
//     accumulates over all the (image, column segment, row block) units it is assigned: split-K over
//     CTAs, one partial per CTA, summed deterministically by wgrad_reduce_kernel;
//   * the input rows live in a shared-memory ring: every row (with its +-p halo columns, zero-filled by
//     TMA = SAME padding) is loaded once per unit and reused by all taps of the group; a tap (ky,kx) is
//     the ring row y+ky-p read at a start address shifted by kx pixels (128 B each);
//   * G rows stream through their own ring; out-of-range columns of the last segment are zero-filled by
//     TMA and so contribute nothing;
//   * the MMA computes the TRANSPOSED tap block D^T[co][ci] (A = G, B = X), because that puts the shifted input on
//     the N side: two vertically adjacent taps (ky, kx), (ky+1, kx) read ring rows r and r+1 at the same column
//     shift, and in the ring those are exactly one "64-channel half" stride apart four times over (half 0 / half 1
//     of row r, half 0 / half 1 of row r+1 -- the descriptor's N-atom stride), so ONE tcgen05.mma of N = 256
//     accumulates both taps: 12 KB of shared-memory operand reads per 2 taps instead of 16, and a paired tap was
//     measured to cost 0.75 of a single one.  A group holds items of 1 or 2 taps; the drain transposes
//     back.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "internal.h"
#include "ptx.cuh"

namespace sr {

namespace {

constexpr int kWgThreads = 256;  // warp0: TMA, warp1: MMA, warp2: TMEM alloc, warps4-7: final TMEM drain
constexpr int kMaxGroups = 8;
constexpr int kMaxRing = 12;
constexpr int kMaxBSlots = 6;
constexpr int kAccFloats = 128 * 128;

struct WgradParams {
  int k, p, NB, H, W;
  int BW, nseg, PWs;   // segment width (multiple of 16), segments per row, ring row pitch (multiple of 8)
  int RB, nrb;         // rows per unit, row blocks per image column
  int NI, ngrp;        // images per K row (narrow images are processed NI at a time), image groups
  int nring, nbslots;
  int ngroups;
  int dbg;              // development switches (SR100_WGRAD_DBG): 1 = no TMA after the first ring fill, 2 = no MMA
  int g_ntaps[kMaxGroups], g_cta0[kMaxGroups], g_ncta[kMaxGroups];
  // items of a group: a single tap (w = 1, 128 TMEM columns) or a vertical tap pair (ky, kx), (ky+1, kx) (w = 2,
  // 256 columns, one N = 256 MMA); col = first 128-column unit of the item in the CTA's TMEM / partial block
  int g_nitems[kMaxGroups];
  signed char it_ky[kMaxGroups][4], it_kx[kMaxGroups][4], it_w[kMaxGroups][4], it_col[kMaxGroups][4];
  signed char tap_group[25], tap_col[25];   // filter tap -> group, 128-column unit (for the reduction)
  float* partial;      // [grid][4][128 ci][128 co]
};

struct __align__(8) WgradBarriers {
  uint64_t a_full[kMaxRing], a_empty[kMaxRing];
  uint64_t b_full[kMaxBSlots], b_empty[kMaxBSlots];
  uint64_t acc_full;
  uint32_t tmem_base, pad;
};

//   [4,6) c format f32, [7,10) a bf16, [10,13) b bf16, [15] a MN-major, [16] b MN-major, N>>3, M>>4
constexpr uint32_t kIdescMN = umma_idesc(1u, 128u, 128u) | (1u << 15) | (1u << 16);
constexpr uint32_t kIdescMN256 = umma_idesc(1u, 128u, 256u) | (1u << 15) | (1u << 16);

}  // namespace

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                const WgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t a_img_bytes = (uint32_t)P.PWs * 128u, b_img_bytes = (uint32_t)P.BW * 128u;  // one image, one half
  const uint32_t a_row_bytes = (uint32_t)P.NI * a_img_bytes * 2u;  // two 64-channel halves of NI x PWs x 128 B
  const uint32_t b_row_bytes = (uint32_t)P.NI * b_img_bytes * 2u;
  uint8_t* a_buf = smem;
  uint8_t* b_buf = a_buf + (size_t)P.nring * a_row_bytes;
  WgradBarriers* bars = reinterpret_cast<WgradBarriers*>(b_buf + (size_t)P.nbslots * b_row_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tap group and rank of this CTA inside it
  int g = 0;
  for (int i = 0; i < P.ngroups; ++i)
    if ((int)blockIdx.x >= P.g_cta0[i]) g = i;
  const int rank = (int)blockIdx.x - P.g_cta0[g];
  const int ncta = P.g_ncta[g];
  const int nitems = P.g_nitems[g];
  int kymin = 99, kymax = -1, ncols = 0;
  for (int j = 0; j < nitems; ++j) {
    kymin = min(kymin, (int)P.it_ky[g][j]);
    kymax = max(kymax, (int)P.it_ky[g][j] + P.it_w[g][j] - 1);
    ncols += P.it_w[g][j];
  }
  const int span = kymax - kymin;  // extra input rows per unit
  const int units = P.ngrp * P.nseg * P.nrb;
  const bool has_work = rank < units;

  if (threadIdx.x == 0) {
    for (int i = 0; i < P.nring; ++i) {
      mbar_init(&bars->a_full[i], 1);
      mbar_init(&bars->a_empty[i], 1);
    }
    for (int i = 0; i < P.nbslots; ++i) {
      mbar_init(&bars->b_full[i], 1);
      mbar_init(&bars->b_empty[i], 1);
    }
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmG);
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  griddep_wait();   // programmatic dependent launch: the set-up above overlapped the predecessor's tail
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch_dependents();
  const uint32_t tmem_base = bars->tmem_base;

  auto decode = [&](int u, int* n, int* x0, int* y0, int* rows) {
    const int rb = u % P.nrb;
    const int q = u / P.nrb;
    const int seg = q % P.nseg;
    *n = (q / P.nseg) * P.NI;  // first image of the group (images beyond NB are zero-filled by TMA)
    *x0 = seg * P.BW;
    *y0 = rb * P.RB;
    *rows = min(P.RB, P.H - *y0);
  };

  if (warp == 0) {
    // ------------------------------------------------ TMA producer: input rows (ring) and gradient rows
    if (lane == 0 && has_work) {
      uint32_t a_slot = 0, a_ph = 0, b_slot = 0, b_ph = 0;  // ring positions (slot + phase parity)
      bool a_wrapped = false, b_wrapped = false;
      for (int u = rank; u < units; u += ncta) {
        int n, x0, y0, rows;
        decode(u, &n, &x0, &y0, &rows);
        int next_in = y0 + kymin - P.p;  // next input row to load
        for (int y = y0; y < y0 + rows; ++y) {
          const int need = y + kymax - P.p;  // newest input row this output row reads
          for (; next_in <= need; ++next_in) {
            mbar_wait(&bars->a_empty[a_slot], a_ph ^ 1u);
            if (SR_DBG(P, 1) && a_wrapped) {
              mbar_arrive(&bars->a_full[a_slot]);
            } else {
              mbar_expect_tx(&bars->a_full[a_slot], a_row_bytes);
              uint8_t* dst = a_buf + (size_t)a_slot * a_row_bytes;
              tma_load_4d(dst, &tmX, &bars->a_full[a_slot], 0, x0 - P.p, next_in, n);
              tma_load_4d(dst + a_row_bytes / 2, &tmX, &bars->a_full[a_slot], 64, x0 - P.p, next_in, n);
            }
            if (++a_slot == (uint32_t)P.nring) {
              a_slot = 0;
              a_ph ^= 1u;
              a_wrapped = true;
            }
          }
          mbar_wait(&bars->b_empty[b_slot], b_ph ^ 1u);
          if (SR_DBG(P, 1) && b_wrapped) {
            mbar_arrive(&bars->b_full[b_slot]);
          } else {
            mbar_expect_tx(&bars->b_full[b_slot], b_row_bytes);
            uint8_t* dst = b_buf + (size_t)b_slot * b_row_bytes;
            tma_load_4d(dst, &tmG, &bars->b_full[b_slot], 0, x0, y, n);
            tma_load_4d(dst + b_row_bytes / 2, &tmG, &bars->b_full[b_slot], 64, x0, y, n);
          }
          if (++b_slot == (uint32_t)P.nbslots) {
            b_slot = 0;
            b_ph ^= 1u;
            b_wrapped = true;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (warp-convergent, elected lane issues)
    // No `if (has_work)` around the issue loop (a CTA without units simply runs zero iterations): under that branch
    // ptxas treated the whole loop as possibly divergent and kept ring positions, descriptors and trip counts in
    // thread registers -- 102 R2UR between the first and the last UTCHMMA, 64 registers; without it the loop lives in
    // uniform registers like the conv kernels' (no R2UR, 45 registers; tests/test_sass_guard.py).
    {
      const bool leader = elect_one();
      // MN-major SW128 descriptors: LBO = bytes between the two 64-channel halves, SBO = 8 pixel rows
      const uint32_t hi = (1024u >> 4) | (1u << 14) | ((uint32_t)SR_LAYOUT_SW128 << 29);
      const uint32_t a_row16 = a_row_bytes >> 4, b_row16 = b_row_bytes >> 4;
      const uint32_t a_img16 = a_img_bytes >> 4, b_img16 = b_img_bytes >> 4;
      const uint32_t a_lo0 = (smem_u32(a_buf) >> 4) | (((a_row_bytes / 2) >> 4) << 16);
      const uint32_t b_lo0 = (smem_u32(b_buf) >> 4) | (((b_row_bytes / 2) >> 4) << 16);
      const int k16n = P.BW >> 4;
      const uint32_t nring = (uint32_t)P.nring, nbs = (uint32_t)P.nbslots;
      // per-item ring-row offset (ky - kymin), start-address shift (kx pixels x 128 B, >> 4), width, TMEM column
      uint32_t t_row[4], t_sh[4], t_w[4], t_col[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jj = j < nitems ? j : 0;
        t_row[j] = (uint32_t)(P.it_ky[g][jj] - kymin);
        t_sh[j] = (uint32_t)P.it_kx[g][jj] * 8u;
        t_w[j] = (uint32_t)P.it_w[g][jj];
        t_col[j] = (uint32_t)P.it_col[g][jj] * 128u;
      }
      uint32_t base = 0;                   // ring slot of the oldest live input row (r = yy)
      uint32_t new_slot = 0, new_ph = 0;   // ring position of the next input row to become full
      uint32_t b_slot = 0, b_ph = 0;
      uint32_t started = 0;
      for (int u = rank; u < units; u += ncta) {
        int n, x0, y0, rows;
        decode(u, &n, &x0, &y0, &rows);
        for (int yy = 0; yy < rows; ++yy) {
          // rows become full in order; output row yy newly needs input row yy+span (rows 0..span at yy == 0)
          for (int r = (yy == 0 ? 0 : span); r <= span; ++r) {
            mbar_wait(&bars->a_full[new_slot], new_ph);
            if (++new_slot == nring) {
              new_slot = 0;
              new_ph ^= 1u;
            }
          }
          mbar_wait(&bars->b_full[b_slot], b_ph);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + b_slot * b_row16;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < nitems) {
              uint32_t slot = base + t_row[j];
              if (slot >= nring) slot -= nring;
              // a pair needs ring rows slot, slot + 1 adjacent in memory; across the ring's wrap it is two N = 128 MMAs
              const bool wide = t_w[j] == 2u && slot + 1u < nring;
              const uint32_t nsub = (t_w[j] == 2u && !wide) ? 2u : 1u;
              for (uint32_t sub = 0; sub < nsub; ++sub) {
                const uint32_t sl = sub ? 0u : slot;
                const uint32_t x_tap = a_lo0 + sl * a_row16 + t_sh[j];
                const uint32_t d = tmem_base + t_col[j] + sub * 128u;
                uint32_t accf = started;
                for (int im = 0; im < P.NI; ++im) {
                  uint32_t x_lo = x_tap + (uint32_t)im * a_img16;
                  uint32_t gl = b_lo + (uint32_t)im * b_img16;
                  for (int s = 0; s < k16n; ++s) {
                    if (leader && !SR_DBG(P, 2)) {
                      const uint64_t gdesc = ((uint64_t)hi << 32) | (uint64_t)gl;    // A: G rows, M = co
                      const uint64_t xdesc = ((uint64_t)hi << 32) | (uint64_t)x_lo;  // B: shifted X rows, N = ci (x2)
                      umma_bf16(d, gdesc, xdesc, wide ? kIdescMN256 : kIdescMN, accf);
                    }
                    accf = 1u;
                    x_lo += 128u;
                    gl += 128u;
                  }
                }
              }
            }
          }
          started = 1u;
          if (leader) {
            umma_commit(&bars->b_empty[b_slot]);
            umma_commit(&bars->a_empty[base]);  // the oldest input row is not read by later output rows
          }
          if (++b_slot == nbs) {
            b_slot = 0;
            b_ph ^= 1u;
          }
          if (++base == nring) base = 0;
        }
        // the last `span` input rows of the unit are dead too
        for (int r = 0; r < span; ++r) {
          if (leader) umma_commit(&bars->a_empty[base]);
          if (++base == nring) base = 0;
        }
      }
      if (has_work) {
        if (leader) umma_commit(&bars->acc_full);
        // wait for the accumulators here (one polling warp), then release the drain warps from their
        // hardware barrier: they sleep instead of polling an mbarrier for the whole kernel
        mbar_wait(&bars->acc_full, 0);
      }
    }
    __syncwarp();
    asm volatile("bar.sync 1, 160;" ::: "memory");
  } else if (warp >= 4) {
    // ------------------------------------------------ drain: TMEM -> partial[cta][tap][ci][co]
    const int ew = warp - 4;
    float* dst = P.partial + (size_t)blockIdx.x * 4 * kAccFloats;
    asm volatile("bar.sync 1, 160;" ::: "memory");
    if (has_work) {
      tc_fence_after();
      // the accumulators hold D^T: TMEM lane = co, column = ci.  A lane writes its co of 32 consecutive ci rows, so
      // every store instruction of the warp covers 32 consecutive floats of one [ci] row of the partial block.
      const int co = ew * 32 + lane;
      for (int j = 0; j < ncols; ++j) {
        float* blk = dst + (size_t)j * kAccFloats + co;
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(j * 128 + cb * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) blk[(size_t)(cb * 32 + e) * 128] = __uint_as_float(v[e]);
        }
      }
    } else {
      for (int j = 0; j < ncols; ++j) {
        float4* row = reinterpret_cast<float4*>(dst + (size_t)j * kAccFloats + (size_t)(ew * 32 + lane) * 128);
        for (int e = 0; e < 32; ++e) row[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dW[tap][ci][co] (= HWIO) = beta * dW + scale * sum over the CTAs of the tap's group of their partials.
__global__ void wgrad_reduce_kernel(const WgradParams P, float scale, float beta, float* __restrict__ dw) {
  griddep_wait();
  griddep_launch_dependents();
  const int total4 = P.k * P.k * kAccFloats / 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
    const int tap = i / (kAccFloats / 4);
    const int e4 = i - tap * (kAccFloats / 4);
    const int g = P.tap_group[tap], j = P.tap_col[tap];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < P.g_ncta[g]; ++c) {
      const float4 v = reinterpret_cast<const float4*>(
          P.partial + ((size_t)(P.g_cta0[g] + c) * 4 + j) * kAccFloats)[e4];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float4* o = reinterpret_cast<float4*>(dw) + i;
    float4 r = make_float4(scale * acc.x, scale * acc.y, scale * acc.z, scale * acc.w);
    if (beta != 0.f) {
      const float4 old = *o;
      r.x = fmaf(beta, old.x, r.x); r.y = fmaf(beta, old.y, r.y);
      r.z = fmaf(beta, old.z, r.z); r.w = fmaf(beta, old.w, r.w);
    }
    *o = r;
  }
}

// =====================================================================================
// Host side
// =====================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled wg_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static int make_row_map(CUtensorMap* tm, const void* ptr, int NB, int H, int W, int box_w, int box_n) {
  PFN_encodeTiled enc = wg_encode_fn();
  if (!enc) return set_error(SR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {128, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
  cuuint64_t strides[3] = {256, (cuuint64_t)W * 256, (cuuint64_t)H * W * 256};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, 1, (cuuint32_t)box_n};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(wgrad row map) failed: %d (W=%d H=%d box=%d)", (int)r, W,
             H, box_w);
    return set_error(SR_ERR_CUDA, msg);
  }
  return SR_OK;
}

struct WgradPlan {
  CUtensorMap tmX, tmG;
  WgradParams P;
  int grid;
  size_t smem_bytes;
  float scale, beta;
  float* dw;
  double flops;
};

static constexpr size_t kWgSmemBudget = 227 * 1024;
static constexpr double kWgCostFloor = 0.0;   // see the CTA split in sr_wgrad_plan_create

}  // namespace sr

using namespace sr;

extern "C" size_t sr_wgrad_workspace_bytes(void) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  return (size_t)sms * 4 * kAccFloats * sizeof(float);
}

extern "C" int sr_wgrad_plan_create(const sr_wgrad_desc* d, sr_wgrad_plan** out) {
  if (!d || !out) return set_error(SR_ERR_INVALID, "sr_wgrad_plan_create: null argument");
  if (!d->x_bf16 || !d->g_bf16 || !d->dw_hwio || !d->workspace)
    return set_error(SR_ERR_INVALID, "sr_wgrad_plan_create: null tensor pointer");
  if (!(d->ksize == 1 || d->ksize == 3 || d->ksize == 5))
    return set_error(SR_ERR_UNSUPPORTED, "wgrad kernel size must be 1, 3 or 5");
  if (d->NB < 1 || d->H < 1 || d->W < 1) return set_error(SR_ERR_INVALID, "empty tensor");
  if (d->workspace_bytes < sr_wgrad_workspace_bytes())
    return set_error(SR_ERR_INVALID, "wgrad workspace too small (see sr_wgrad_workspace_bytes)");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;

  WgradPlan* pl = new (std::nothrow) WgradPlan();
  if (!pl) return set_error(SR_ERR_NOMEM, "out of host memory");
  memset(pl, 0, sizeof *pl);
  WgradParams& P = pl->P;
  P.k = d->ksize;
  P.p = (d->ksize - 1) / 2;
  P.NB = d->NB;
  P.H = d->H;
  P.W = d->W;
  const int ntaps = P.k * P.k;

  // Tap groups: <= 4 taps = 512 TMEM columns each, sizes as even as possible (a CTA's share of the SMs follows its
  // group's tap count).  Vertical tap pairs become one item (one N = 256 MMA per K step); the tables below tile
  // the 3x3 / 5x5 taps with such pairs plus single taps so that every group keeps 3-4 taps and spans <= 3 input
  // rows.  SR100_WGRAD_VPAIR=0: single taps only (groups of consecutive taps, the earlier scheme).
  struct Item { int ky, kx, w; };
  static const Item k5_items[7][4] = {
      {{0, 0, 2}, {0, 1, 2}, {-1, 0, 0}, {-1, 0, 0}}, {{0, 2, 2}, {0, 3, 2}, {-1, 0, 0}, {-1, 0, 0}},
      {{2, 0, 2}, {2, 1, 2}, {-1, 0, 0}, {-1, 0, 0}}, {{2, 2, 2}, {2, 3, 2}, {-1, 0, 0}, {-1, 0, 0}},
      {{0, 4, 1}, {1, 4, 2}, {-1, 0, 0}, {-1, 0, 0}}, {{3, 4, 2}, {4, 3, 1}, {-1, 0, 0}, {-1, 0, 0}},
      {{4, 0, 1}, {4, 1, 1}, {4, 2, 1}, {-1, 0, 0}}};
  static const Item k3_items[3][4] = {{{0, 0, 2}, {2, 0, 1}, {-1, 0, 0}, {-1, 0, 0}},
                                      {{0, 1, 2}, {2, 1, 1}, {-1, 0, 0}, {-1, 0, 0}},
                                      {{0, 2, 2}, {2, 2, 1}, {-1, 0, 0}, {-1, 0, 0}}};
  const char* vp_env = dev_getenv("SR100_WGRAD_VPAIR");
  const bool vpair = !(vp_env && atoi(vp_env) == 0) && (P.k == 3 || P.k == 5);
  P.ngroups = (ntaps + 3) / 4;
  int max_span = 0;
  {
    int t = 0;
    for (int g = 0; g < P.ngroups; ++g) {
      int kymin = 99, kymax = -1, col = 0, nt = 0, ni = 0;
      if (vpair) {
        const Item* items = P.k == 5 ? k5_items[g] : k3_items[g];
        for (int j = 0; j < 4 && items[j].ky >= 0; ++j, ++ni) {
          P.it_ky[g][j] = (signed char)items[j].ky;
          P.it_kx[g][j] = (signed char)items[j].kx;
          P.it_w[g][j] = (signed char)items[j].w;
          P.it_col[g][j] = (signed char)col;
          for (int q = 0; q < items[j].w; ++q) {
            const int tap = (items[j].ky + q) * P.k + items[j].kx;
            P.tap_group[tap] = (signed char)g;
            P.tap_col[tap] = (signed char)(col + q);
          }
          kymin = std::min(kymin, items[j].ky);
          kymax = std::max(kymax, items[j].ky + items[j].w - 1);
          col += items[j].w;
          nt += items[j].w;
        }
      } else {
        const int left = ntaps - t, gl = P.ngroups - g;
        const int sz = (left + gl - 1) / gl;
        for (int j = 0; j < sz; ++j, ++t, ++ni) {
          P.it_ky[g][j] = (signed char)(t / P.k);
          P.it_kx[g][j] = (signed char)(t % P.k);
          P.it_w[g][j] = 1;
          P.it_col[g][j] = (signed char)j;
          P.tap_group[t] = (signed char)g;
          P.tap_col[t] = (signed char)j;
          kymin = std::min(kymin, t / P.k);
          kymax = std::max(kymax, t / P.k);
        }
        nt = sz;
      }
      P.g_nitems[g] = ni;
      P.g_ntaps[g] = nt;
      max_span = std::max(max_span, kymax - kymin);
    }
  }

  // geometry: segment width (multiple of 16, <= 128) and ring depth within the shared-memory budget
  bool ok = false;
  for (int maxbw = 128; maxbw >= 16 && !ok; maxbw -= 16) {
    const int nseg = (d->W + maxbw - 1) / maxbw;
    const int bw = (((d->W + nseg - 1) / nseg) + 15) & ~15;
    if (bw > maxbw) continue;
    const int pws = (bw + 2 * P.p + 7) & ~7;
    if (pws > 256) continue;
    const size_t fixed = 1024 + sizeof(WgradBarriers) + 64;
    const int min_ring = max_span + 2, min_b = 2;
    // narrow images: NI images share one K row (one barrier round trip feeds NI x BW/16 MMAs per tap)
    int ni = nseg == 1 ? std::max(1, std::min(std::min(4, d->NB), 128 / bw)) : 1;
    while (ni > 1 && fixed + (size_t)(min_ring + 1) * ni * pws * 256 + (size_t)(min_b + 1) * ni * bw * 256 > kWgSmemBudget)
      --ni;
    const size_t a_row = (size_t)ni * pws * 256, b_row = (size_t)ni * bw * 256;
    if (fixed + min_ring * a_row + min_b * b_row > kWgSmemBudget) continue;
    int nring = min_ring, nb = min_b;
    // grow both rings alternately while they fit
    for (;;) {
      bool grew = false;
      if (nb < kMaxBSlots && nb < 4 && fixed + nring * a_row + (nb + 1) * b_row <= kWgSmemBudget) {
        ++nb;
        grew = true;
      }
      if (nring < kMaxRing && nring < max_span + 5 && fixed + (nring + 1) * a_row + nb * b_row <= kWgSmemBudget) {
        ++nring;
        grew = true;
      }
      if (!grew) break;
    }
    if (const char* e = dev_getenv("SR100_WGRAD_RING")) {   // development override: "<ring>,<g slots>"
      int r_ = 0, b_ = 0;
      if (sscanf(e, "%d,%d", &r_, &b_) == 2 && r_ >= min_ring && r_ <= kMaxRing && b_ >= min_b && b_ <= kMaxBSlots &&
          fixed + r_ * a_row + b_ * b_row <= kWgSmemBudget) {
        nring = r_;
        nb = b_;
      }
    }
    P.BW = bw;
    P.NI = ni;
    P.ngrp = (d->NB + ni - 1) / ni;
    P.nseg = nseg;
    P.PWs = pws;
    P.nring = nring;
    P.nbslots = nb;
    pl->smem_bytes = fixed + nring * a_row + nb * b_row;
    ok = true;
  }
  if (!ok) {
    delete pl;
    return set_error(SR_ERR_UNSUPPORTED, "no wgrad geometry fits shared memory");
  }

  // CTAs per group proportional to the group's cost: a tap of a vertical pair is cheaper than a single tap (one
  // N = 256 MMA for two taps reads 12 KB of shared memory instead of 16: 0.75 of a single tap, and measured so)
  double cost[kMaxGroups], total_cost = 0;
  {
    const char* e = dev_getenv("SR100_WGRAD_PCOST");
    const double pair_cost = e ? atof(e) : 1.5;    // cost of a pair item in single-tap units (measured optimum)
    // every group streams the same X and G rows whatever its tap count: a floor on the cost is the part of a group's
    // time that is operand ingest rather than MMA work (floor >= the largest group cost: equal CTAs per group)
    const char* ef = dev_getenv("SR100_WGRAD_COSTFLOOR");
    const double cost_floor = ef ? atof(ef) : kWgCostFloor;
    for (int g = 0; g < P.ngroups; ++g) {
      cost[g] = 0;
      for (int j = 0; j < P.g_nitems[g]; ++j) cost[g] += P.it_w[g][j] == 2 ? pair_cost : 1.0;
      cost[g] = std::max(cost[g], cost_floor);
      total_cost += cost[g];
    }
  }
  int assigned = 0;
  for (int g = 0; g < P.ngroups; ++g) {
    int n = (int)((double)sms * cost[g] / total_cost);
    if (n < 1) n = 1;
    P.g_ncta[g] = n;
    assigned += n;
  }
  while (assigned < sms) {   // hand out the remainder to the group with the highest cost per CTA
    int best = 0;
    for (int g = 1; g < P.ngroups; ++g)
      if (cost[g] / P.g_ncta[g] > cost[best] / P.g_ncta[best]) best = g;
    ++P.g_ncta[best];
    ++assigned;
  }
  while (assigned > sms) {
    for (int g = P.ngroups - 1; g >= 0 && assigned > sms; --g)
      if (P.g_ncta[g] > 1) {
        --P.g_ncta[g];
        --assigned;
      }
  }
  int max_ncta = 0;
  for (int g = 0; g < P.ngroups; ++g) max_ncta = std::max(max_ncta, P.g_ncta[g]);
  P.RB = P.H;
  P.nrb = 1;
  while ((long long)P.ngrp * P.nseg * P.nrb < 12LL * max_ncta && P.RB > 8) {
    P.RB = (P.RB + 1) / 2;
    P.nrb = (P.H + P.RB - 1) / P.RB;
  }
  {
    int c = 0;
    for (int g = 0; g < P.ngroups; ++g) {
      P.g_cta0[g] = c;
      c += P.g_ncta[g];
    }
    pl->grid = c;
  }
  P.partial = reinterpret_cast<float*>(d->workspace);
  {
    const char* e = dev_getenv("SR100_WGRAD_DBG");
    P.dbg = e ? atoi(e) : 0;
  }
  pl->scale = d->scale;
  pl->beta = d->accumulate ? 1.f : 0.f;
  pl->dw = d->dw_hwio;
  pl->flops = 2.0 * (double)d->NB * d->H * d->W * ntaps * 128.0 * 128.0;
  int rc = make_row_map(&pl->tmX, d->x_bf16, d->NB, d->H, d->W, P.PWs, P.NI);
  if (rc == SR_OK) rc = make_row_map(&pl->tmG, d->g_bf16, d->NB, d->H, d->W, P.BW, P.NI);
  if (rc != SR_OK) {
    delete pl;
    return rc;
  }
  *out = reinterpret_cast<sr_wgrad_plan*>(pl);
  return SR_OK;
}

extern "C" int sr_wgrad_plan_run(sr_wgrad_plan* plan, void* stream) {
  if (!plan) return set_error(SR_ERR_INVALID, "sr_wgrad_plan_run: null plan");
  const WgradPlan* pl = reinterpret_cast<const WgradPlan*>(plan);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static unsigned long long attr_done = 0;
  if (int rc = ensure_dynamic_smem(wgrad_tc_kernel, (int)kWgSmemBudget, &attr_done, "cudaFuncSetAttribute(wgrad_tc_kernel)"))
    return rc;
  cudaError_t e = launch_pdl(wgrad_tc_kernel, pl->grid, kWgThreads, pl->smem_bytes, st, pl->tmX, pl->tmG, pl->P);
  if (e != cudaSuccess) return set_cuda_error(e, "wgrad_tc_kernel launch");
  e = launch_pdl(wgrad_reduce_kernel, 296u, 256u, 0, st, pl->P, pl->scale, pl->beta, pl->dw);
  if (e != cudaSuccess) return set_cuda_error(e, "wgrad_reduce_kernel launch");
  return SR_OK;
}

extern "C" void sr_wgrad_plan_destroy(sr_wgrad_plan* plan) { delete reinterpret_cast<WgradPlan*>(plan); }

extern "C" int sr_wgrad_plan_info(const sr_wgrad_plan* plan, sr_wgrad_plan_info_t* info) {
  if (!plan || !info) return set_error(SR_ERR_INVALID, "sr_wgrad_plan_info: null argument");
  const WgradPlan* pl = reinterpret_cast<const WgradPlan*>(plan);
  info->flops = pl->flops;
  info->grid = pl->grid;
  info->smem_bytes = (int)pl->smem_bytes;
  info->seg_width = pl->P.BW;
  info->nseg = pl->P.nseg;
  info->ring_rows = pl->P.nring;
  info->g_slots = pl->P.nbslots;
  info->tap_groups = pl->P.ngroups;
  info->rows_per_unit = pl->P.RB;
  info->images_per_row = pl->P.NI;
  return SR_OK;
}
