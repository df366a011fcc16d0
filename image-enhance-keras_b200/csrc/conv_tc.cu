// tcgen05 implicit-GEMM convolution kernel + host-side plan (geometry, TMA descriptors, launch).
// See conv_tc.cuh for the formulation.  Reference semantics: Keras Conv2D(padding='same'),
// models.py:1177-1199, 1231-1270 (cross-correlation, HWIO kernels, bias, optional ReLU) and the
// residual algebra of _residual_block_light / _residual_block_light53 (0.1 / 0.9 scalar_mul,
// models.py:977-986), which is fused here as out = act(alpha*(acc+bias) + beta*res).
#include "conv_tc.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "internal.h"
#include "ptx.cuh"

namespace sr {

namespace {

constexpr int kMaxWStages = 16;
constexpr int kTapsPerStage = 2;  // taps per weight stage: halves the barrier round trips per MMA

struct __align__(8) ConvBarriers {
  uint64_t w_full[kMaxWStages];
  uint64_t w_empty[kMaxWStages];
  uint64_t a_full[4];
  uint64_t a_empty[4];
  uint64_t tmem_full[4];
  uint64_t tmem_empty[4];
  uint32_t tmem_base;
  uint32_t pad;
};

// position in an mbarrier ring: slot index + phase parity (avoids a runtime modulo per stage)
struct Ring {
  uint32_t slot, phase, n;
  __device__ __forceinline__ void advance() {
    if (++slot == n) {
      slot = 0;
      phase ^= 1u;
    }
  }
};

struct TileCoord {
  int n, seg_x0, f0, r_lo, off0;
};

template <int T>
__device__ __forceinline__ TileCoord decode_tile(const ConvKernelParams& P, int t) {
  const int per_img = P.nseg * P.tiles_per_seg;
  TileCoord c;
  c.n = t / per_img;
  const int r = t - c.n * per_img;
  const int seg = r / P.tiles_per_seg;
  const int ti = r - seg * P.tiles_per_seg;
  c.seg_x0 = seg * P.BW;
  c.f0 = P.p * P.PWs + P.p + ti * T;
  c.r_lo = (ti * T) / P.PWs;  // first strip row, in padded-row units (row 0 = image row -p)
  c.off0 = c.f0 - c.r_lo * P.PWs;
  return c;
}

// Generic staged epilogue (any combination of residual / mask, all selected at run time) for one 128x128 accumulator: phase 1 drains TMEM (lane = pixel) into this warp's
// 32-row x 256 B shared-memory tile as bf16 (16-byte chunks XOR-swizzled by row, conflict-free both
// ways); phase 2 re-reads it with 16 lanes per pixel so that every global access of the residual,
// mask and outputs is a fully coalesced 512 B (bf16) / 1 KB (fp32) warp transaction.
__device__ __forceinline__ void epilogue_staged_acc_generic(const ConvKernelParams& P, const TileCoord& c,
                                                    uint32_t t_acc, int f_base, uint8_t* stage,
                                                    const float* s_bias, int lane, bool live = true) {
  // ---- phase 1: this lane's pixel row
  const int f = f_base + lane;
  const int fr = f / P.PWs;
  const int yy = fr - P.p;
  const int xx = f - fr * P.PWs - P.p;
  const bool valid = live && (f - (P.p * P.PWs + P.p) < P.f_len) && yy < P.Hc && xx >= 0 &&
                     xx < P.BW && (c.seg_x0 + xx) < P.Wc;
  const int my_pix = valid ? (int)(((size_t)c.n * P.H + yy) * P.W + c.seg_x0 + xx) : -1;
  uint8_t* my_row = stage + lane * 256;
  const int sw = lane & 7;
  // two TMEM loads in flight per wait (64 columns): halves the exposed tcgen05.ld latency
#pragma unroll 1
  for (int cb2 = 0; cb2 < 2; ++cb2) {
    uint32_t va[32], vb[32];
    tmem_ld32(t_acc + cb2 * 64, va);
    tmem_ld32(t_acc + cb2 * 64 + 32, vb);
    tmem_ld_wait();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = q * 8 + e * 2;
          const float f0 = __uint_as_float(hh ? vb[j] : va[j]), f1 = __uint_as_float(hh ? vb[j + 1] : va[j + 1]);
          const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
          w[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(my_row + ((((cb2 * 2 + hh) * 4 + q) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  __syncwarp();
  // ---- phase 2: 16 lanes per pixel, 2 pixels per warp instruction
  const int half = lane >> 4, q = lane & 15;
  float bs[8];  // alpha * bias of this lane's 8 channels
#pragma unroll
  for (int e = 0; e < 8; ++e) bs[e] = P.alpha * s_bias[q * 8 + e];
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int r = 2 * i + half;
    const int pix = __shfl_sync(0xffffffffu, my_pix, r);
    if (pix < 0) continue;
    const uint4 sv = *reinterpret_cast<const uint4*>(stage + r * 256 + ((q ^ (r & 7)) << 4));
    const uint32_t w[4] = {sv.x, sv.y, sv.z, sv.w};
    float o[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      o[2 * e] = fmaf(P.alpha, __uint_as_float(w[e] << 16), bs[2 * e]);
      o[2 * e + 1] = fmaf(P.alpha, __uint_as_float(w[e] & 0xFFFF0000u), bs[2 * e + 1]);
    }
    const size_t off = (size_t)pix * 128 + q * 8;
    if (P.res_f32) {
      const float4 r0 = *reinterpret_cast<const float4*>(P.res_f32 + off);
      const float4 r1 = *reinterpret_cast<const float4*>(P.res_f32 + off + 4);
      o[0] = fmaf(P.beta, r0.x, o[0]); o[1] = fmaf(P.beta, r0.y, o[1]);
      o[2] = fmaf(P.beta, r0.z, o[2]); o[3] = fmaf(P.beta, r0.w, o[3]);
      o[4] = fmaf(P.beta, r1.x, o[4]); o[5] = fmaf(P.beta, r1.y, o[5]);
      o[6] = fmaf(P.beta, r1.z, o[6]); o[7] = fmaf(P.beta, r1.w, o[7]);
    } else if (P.res_bf16) {
      const uint4 rv = *reinterpret_cast<const uint4*>(P.res_bf16 + off);
      const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[2 * e] = fmaf(P.beta, __uint_as_float(rw[e] << 16), o[2 * e]);
        o[2 * e + 1] = fmaf(P.beta, __uint_as_float(rw[e] & 0xFFFF0000u), o[2 * e + 1]);
      }
    }
    if (P.relu == 2) {  // keras LeakyReLU(alpha): x >= 0 ? x : alpha * x
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = o[e] >= 0.f ? o[e] : P.neg_slope * o[e];
    } else if (P.relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], 0.f);
    }
    if (P.relu_mask_bf16) {
      const uint4 mv = *reinterpret_cast<const uint4*>(P.relu_mask_bf16 + off);
      const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (!(__uint_as_float(mw[e] << 16) > 0.f)) o[2 * e] = P.mask_slope == 0.f ? 0.f : o[2 * e] * P.mask_slope;
        if (!(__uint_as_float(mw[e] & 0xFFFF0000u) > 0.f))
          o[2 * e + 1] = P.mask_slope == 0.f ? 0.f : o[2 * e + 1] * P.mask_slope;
      }
    }
    if (P.out_f32) {
      *reinterpret_cast<float4*>(P.out_f32 + off) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(P.out_f32 + off + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (P.out_bf16) {
      uint32_t pw[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
        pw[e] = *reinterpret_cast<const uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(P.out_bf16 + off) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
    }
  }
  __syncwarp();
}

// Specialised staged epilogue (EPI = which single global operand the epilogue reads: 1 fp32 residual, 2 bf16
// residual, 3 ReLU mask; the host picks the generic kernel for anything else) for one 128x128 accumulator.
// The operand of 8 pixel pairs is requested as one batch BEFORE the TMEM drain / before the pass that uses it, so
// one DRAM latency is exposed per batch instead of one per pixel pair; the two passes share one (rolled) body to
// keep the kernel small (instruction-cache footprint decides the short-MMA-phase launches).
// CS: the launch also produces the column sums of its bf16 output (training: the bias gradient of the layer whose
// output gradient this launch writes, sr_conv_desc.colsum_f32) -- every lane adds the rounded values it stores to
// eight running sums `cs` (its eight channels), reduced once at the end of the kernel.
template <int EPI, bool CS = false>
__device__ __forceinline__ void epilogue_staged_acc(const ConvKernelParams& P, const TileCoord& c,
                                                    uint32_t t_acc, int f_base, uint8_t* stage,
                                                    const float* s_bias, int lane, bool live, float (&cs)[8]) {
  const int f = f_base + lane;
  const int fr = f / P.PWs;
  const int yy = fr - P.p;
  const int xx = f - fr * P.PWs - P.p;
  const bool valid = live && (f - (P.p * P.PWs + P.p) < P.f_len) && yy < P.Hc && xx >= 0 &&
                     xx < P.BW && (c.seg_x0 + xx) < P.Wc;
  const int my_pix = valid ? (int)(((size_t)c.n * P.H + yy) * P.W + c.seg_x0 + xx) : -1;
  const int half = lane >> 4, q = lane & 15;
  constexpr int NPRE = EPI == 1 ? 16 : 8;
  uint4 pre[NPRE];
  auto prefetch = [&](int pass) {
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      const int pix = __shfl_sync(0xffffffffu, my_pix, 2 * (pass * 8 + ii) + half);
      if constexpr (EPI == 1) {
        pre[2 * ii] = pre[2 * ii + 1] = make_uint4(0u, 0u, 0u, 0u);
        if (pix >= 0) {
          const uint4* rp = reinterpret_cast<const uint4*>(P.res_f32 + (size_t)pix * 128 + q * 8);
          pre[2 * ii] = rp[0];
          pre[2 * ii + 1] = rp[1];
        }
      } else {
        const __nv_bfloat16* src = EPI == 2 ? P.res_bf16 : P.relu_mask_bf16;
        pre[ii] = make_uint4(0u, 0u, 0u, 0u);
        if (pix >= 0) pre[ii] = *reinterpret_cast<const uint4*>(src + (size_t)pix * 128 + q * 8);
      }
    }
  };
  prefetch(0);
  uint8_t* my_row = stage + lane * 256;
  const int sw = lane & 7;
  // ---- phase 1: drain TMEM (lane = pixel), two loads in flight per wait
#pragma unroll 1
  for (int cb2 = 0; cb2 < 2; ++cb2) {
    uint32_t va[32], vb[32];
    tmem_ld32(t_acc + cb2 * 64, va);
    tmem_ld32(t_acc + cb2 * 64 + 32, vb);
    tmem_ld_wait();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = qq * 8 + e * 2;
          const float f0 = __uint_as_float(hh ? vb[j] : va[j]), f1 = __uint_as_float(hh ? vb[j + 1] : va[j + 1]);
          const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
          w[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(my_row + ((((cb2 * 2 + hh) * 4 + qq) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  __syncwarp();
  // ---- phase 2: 16 lanes per pixel, 2 pixels per warp instruction
  float bs[8];  // alpha * bias of this lane's 8 channels
#pragma unroll
  for (int e = 0; e < 8; ++e) bs[e] = P.alpha * s_bias[q * 8 + e];
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1) prefetch(1);
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
      const int r = 2 * (pass * 8 + ii) + half;
      const int pix = __shfl_sync(0xffffffffu, my_pix, r);
      if (pix < 0) continue;
      const uint4 sv = *reinterpret_cast<const uint4*>(stage + r * 256 + ((q ^ (r & 7)) << 4));
      const uint32_t w[4] = {sv.x, sv.y, sv.z, sv.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[2 * e] = fmaf(P.alpha, __uint_as_float(w[e] << 16), bs[2 * e]);
        o[2 * e + 1] = fmaf(P.alpha, __uint_as_float(w[e] & 0xFFFF0000u), bs[2 * e + 1]);
      }
      const size_t off = (size_t)pix * 128 + q * 8;
      if constexpr (EPI == 1) {
        const uint4 r0 = pre[2 * ii], r1 = pre[2 * ii + 1];
        o[0] = fmaf(P.beta, __uint_as_float(r0.x), o[0]); o[1] = fmaf(P.beta, __uint_as_float(r0.y), o[1]);
        o[2] = fmaf(P.beta, __uint_as_float(r0.z), o[2]); o[3] = fmaf(P.beta, __uint_as_float(r0.w), o[3]);
        o[4] = fmaf(P.beta, __uint_as_float(r1.x), o[4]); o[5] = fmaf(P.beta, __uint_as_float(r1.y), o[5]);
        o[6] = fmaf(P.beta, __uint_as_float(r1.z), o[6]); o[7] = fmaf(P.beta, __uint_as_float(r1.w), o[7]);
      } else if constexpr (EPI == 2) {
        const uint4 rv = pre[ii];
        const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          o[2 * e] = fmaf(P.beta, __uint_as_float(rw[e] << 16), o[2 * e]);
          o[2 * e + 1] = fmaf(P.beta, __uint_as_float(rw[e] & 0xFFFF0000u), o[2 * e + 1]);
        }
      }
      if (P.relu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], 0.f);
      }
      if constexpr (EPI == 3) {
        const uint4 mv = pre[ii];
        const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!(__uint_as_float(mw[e] << 16) > 0.f)) o[2 * e] = 0.f;
          if (!(__uint_as_float(mw[e] & 0xFFFF0000u) > 0.f)) o[2 * e + 1] = 0.f;
        }
      }
      if (P.out_f32) {
        *reinterpret_cast<float4*>(P.out_f32 + off) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(P.out_f32 + off + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
      if (P.out_bf16) {
        uint32_t pw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
          pw[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(P.out_bf16 + off) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
        if constexpr (CS) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            cs[2 * e] += __uint_as_float(pw[e] << 16);
            cs[2 * e + 1] += __uint_as_float(pw[e] & 0xFFFF0000u);
          }
        }
      }
    }
  }
  __syncwarp();
}

// Sub-pixel epilogue (EPI = 4): conv + bias (+ReLU) whose fp32 result is stored straight at its depth-to-space
// position (keras_subpixel.py:64-84, advanced.py:87-129,195-196) -- the shuffle is only a store address.
// Staged so that the stores are coalesced: this variant's staging rows are 512 B (all 128 fp32 channels of a
// pixel; kShuffleStageRow), phase 1 drains the accumulator (lane = pixel) into them as finished fp32 values;
// phase 2 walks the OUTPUT: for each sub-row ry the 32 pixels of the warp own 32 runs of r*C consecutive floats
// ([rx][c]; adjacent pixels = adjacent runs), so consecutive lanes take consecutive output words -- 128-bit
// stores when r*C % 4 == 0, 512 B per warp instruction -- and fetch the values from the staging tile through a
// [ry][rx*C+c] -> channel table built once per CTA.  (The first version stored one channel of 32 pixels per
// instruction -- 32 sectors for 128 useful bytes -- and was 4-6x slower than not fusing at all.)
constexpr int kShuffleStageRow = 512;
constexpr int kShuffleStageBytes = 4 * 32 * kShuffleStageRow + 128 * 4;   // 4 warps + the channel table

__device__ __forceinline__ void shuffle_table_init(const ConvKernelParams& P, int* tbl, int tid, int nthreads) {
  const int r_ = P.shuffle_r, C_ = P.shuffle_C, RC = r_ * C_, rr = r_ * r_;
  for (int i = tid; i < P.cout; i += nthreads) {
    const int ry = i / RC, k = i - ry * RC, rx = k / C_, cc = k - rx * C_;
    tbl[i] = P.shuffle_order == 0 ? cc * rr + rx * r_ + ry
           : P.shuffle_order == 1 ? cc * rr + ry * r_ + rx : (ry * r_ + rx) * C_ + cc;
  }
}

__device__ __noinline__ void epilogue_shuffle_acc(const ConvKernelParams& P, const TileCoord& c, uint32_t t_acc,
                                                  int f_base, uint8_t* stage, const int* tbl, const float* s_bias,
                                                  int lane, bool live) {
  const int f = f_base + lane;
  const int fr = f / P.PWs;
  const int yy = fr - P.p;
  const int xx = f - fr * P.PWs - P.p;
  const bool valid = live && (f - (P.p * P.PWs + P.p) < P.f_len) && yy < P.Hc && xx >= 0 &&
                     xx < P.BW && (c.seg_x0 + xx) < P.Wc;
  const int r_ = P.shuffle_r, RC = r_ * P.shuffle_C;
  const int my_row = valid ? c.n * P.H + yy : -1;   // image row index n*H + y of this lane's pixel
  const int my_x = c.seg_x0 + xx;
  const size_t orow = (size_t)P.W * RC;             // floats per output row
  uint8_t* my_srow = stage + lane * kShuffleStageRow;
  const int sw = lane & 7;
  // ---- phase 1: same rounding as the staged epilogues (accumulator -> bf16, then alpha / bias / ReLU in fp32)
  const float lo = P.relu ? 0.f : -3.4e38f;
#pragma unroll 1
  for (int cb = 0; cb < 4; ++cb) {
    if (cb * 32 >= P.cout) break;
    uint32_t v[32];
    tmem_ld32(t_acc + cb * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        o[e] = fmaxf(fmaf(P.alpha, __bfloat162float(__float2bfloat16_rn(__uint_as_float(v[4 * j + e]))),
                          P.alpha * s_bias[cb * 32 + 4 * j + e]), lo);
      *reinterpret_cast<float4*>(my_srow + (((cb * 8 + j) ^ sw) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncwarp();
  // ---- phase 2
  auto fetch = [&](uint32_t p, int ch) {
    return *reinterpret_cast<const float*>(stage + p * kShuffleStageRow + ((((uint32_t)ch >> 2) ^ (p & 7u)) << 4) +
                                           ((ch & 3) << 2));
  };
  if ((RC & 3) == 0) {
    const uint32_t G = (uint32_t)RC >> 2;            // float4 groups per pixel run
    const uint32_t q32 = 32u / G, r32 = 32u % G;
    const uint32_t p0 = (uint32_t)lane / G, g0 = (uint32_t)lane % G;
#pragma unroll 1
    for (int ry = 0; ry < r_; ++ry) {
      const int* t = tbl + ry * RC;
      uint32_t p = p0, g = g0;
#pragma unroll 1
      for (uint32_t it = 0; it < G; ++it) {          // 32*G groups of this sub-row, 32 per iteration
        const int prow = __shfl_sync(0xffffffffu, my_row, (int)p);
        const int px = __shfl_sync(0xffffffffu, my_x, (int)p);
        if (prow >= 0) {
          const int4 ch = *reinterpret_cast<const int4*>(t + 4 * g);
          const float4 o = make_float4(fetch(p, ch.x), fetch(p, ch.y), fetch(p, ch.z), fetch(p, ch.w));
          *reinterpret_cast<float4*>(P.out_f32 + ((size_t)prow * r_ + ry) * orow + (size_t)px * RC + 4 * g) = o;
        }
        p += q32;
        g += r32;
        if (g >= G) { g -= G; ++p; }
      }
    }
  } else {
    const uint32_t q32 = 32u / (uint32_t)RC, r32 = 32u % (uint32_t)RC;
    const uint32_t p0 = (uint32_t)lane / (uint32_t)RC, k0 = (uint32_t)lane % (uint32_t)RC;
#pragma unroll 1
    for (int ry = 0; ry < r_; ++ry) {
      const int* t = tbl + ry * RC;
      uint32_t p = p0, k = k0;
#pragma unroll 1
      for (int it = 0; it < RC; ++it) {
        const int prow = __shfl_sync(0xffffffffu, my_row, (int)p);
        const int px = __shfl_sync(0xffffffffu, my_x, (int)p);
        if (prow >= 0) P.out_f32[((size_t)prow * r_ + ry) * orow + (size_t)px * RC + k] = fetch(p, t[k]);
        p += q32;
        k += r32;
        if (k >= (uint32_t)RC) { k -= RC; ++p; }
      }
    }
  }
  __syncwarp();  // the next accumulator overwrites the staging rows
}

// Lean epilogue for the commonest launch (conv + bias + ReLU -> bf16, no residual / mask / fp32 copy): small code
// keeps the whole kernel inside the instruction cache, which is what bounds the 3x3 convs (their MMA phase per
// tile is too short to hide a slow epilogue).
__device__ __forceinline__ void epilogue_staged_acc_plain(const ConvKernelParams& P, const TileCoord& c,
                                                          uint32_t t_acc, int f_base, uint8_t* stage,
                                                          const float* s_bias, int lane, bool live) {
  const int f = f_base + lane;
  const int fr = f / P.PWs;
  const int yy = fr - P.p;
  const int xx = f - fr * P.PWs - P.p;
  const bool valid = live && (f - (P.p * P.PWs + P.p) < P.f_len) && yy < P.Hc && xx >= 0 &&
                     xx < P.BW && (c.seg_x0 + xx) < P.Wc;
  const int my_pix = valid ? (int)(((size_t)c.n * P.H + yy) * P.W + c.seg_x0 + xx) : -1;
  uint8_t* my_row = stage + lane * 256;
  const int sw = lane & 7;
#pragma unroll 1
  for (int cb = 0; cb < 4; ++cb) {
    uint32_t v[32];
    tmem_ld32(t_acc + cb * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = q * 8 + e * 2;
        const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
        w[e] = *reinterpret_cast<const uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(my_row + (((cb * 4 + q) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  __syncwarp();
  const int half = lane >> 4, q = lane & 15;
  float bs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) bs[e] = P.alpha * s_bias[q * 8 + e];
  const float lo = P.relu ? 0.f : -3.4e38f;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int r = 2 * i + half;
    const int pix = __shfl_sync(0xffffffffu, my_pix, r);
    if (pix < 0) continue;
    const uint4 sv = *reinterpret_cast<const uint4*>(stage + r * 256 + ((q ^ (r & 7)) << 4));
    const uint32_t w[4] = {sv.x, sv.y, sv.z, sv.w};
    uint32_t pw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float o0 = fmaxf(fmaf(P.alpha, __uint_as_float(w[e] << 16), bs[2 * e]), lo);
      const float o1 = fmaxf(fmaf(P.alpha, __uint_as_float(w[e] & 0xFFFF0000u), bs[2 * e + 1]), lo);
      const __nv_bfloat162 h = __floats2bfloat162_rn(o0, o1);
      pw[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(P.out_bf16 + (size_t)pix * 128 + q * 8) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
  }
  __syncwarp();
}

// tf32-mode epilogue (EPI = 5): the accumulator stays fp32 all the way.  A 256 B staging row holds 64 fp32
// channels, so one 128x128 accumulator goes through the warp's staging tile in two halves; phase 2 is the same
// 16-lanes-per-pixel pass (256 B of contiguous fp32 per pixel and half).  out_f32 = the unrounded result (the
// residual stream), out_tf32 = the same value rounded to tf32 (what the next conv's MMAs read).
__device__ __forceinline__ void epilogue_staged_acc_tf32(const ConvKernelParams& P, const TileCoord& c,
                                                         uint32_t t_acc, int f_base, uint8_t* stage,
                                                         const float* s_bias, int lane, bool live) {
  const int f = f_base + lane;
  const int fr = f / P.PWs;
  const int yy = fr - P.p;
  const int xx = f - fr * P.PWs - P.p;
  const bool valid = live && (f - (P.p * P.PWs + P.p) < P.f_len) && yy < P.Hc && xx >= 0 &&
                     xx < P.BW && (c.seg_x0 + xx) < P.Wc;
  const int my_pix = valid ? (int)(((size_t)c.n * P.H + yy) * P.W + c.seg_x0 + xx) : -1;
  uint8_t* my_row = stage + lane * 256;
  const int sw = lane & 7;
  const int half = lane >> 4, q = lane & 15;
#pragma unroll 1
  for (int hb = 0; hb < 2; ++hb) {
    {
      uint32_t va[32], vb[32];
      tmem_ld32(t_acc + hb * 64, va);
      tmem_ld32(t_acc + hb * 64 + 32, vb);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<uint4*>(my_row + ((j ^ sw) << 4)) = make_uint4(va[4 * j], va[4 * j + 1], va[4 * j + 2], va[4 * j + 3]);
        *reinterpret_cast<uint4*>(my_row + (((8 + j) ^ sw) << 4)) = make_uint4(vb[4 * j], vb[4 * j + 1], vb[4 * j + 2], vb[4 * j + 3]);
      }
    }
    __syncwarp();
    float bs[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) bs[e] = P.alpha * s_bias[hb * 64 + q * 4 + e];
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int r = 2 * i + half;
      const int pix = __shfl_sync(0xffffffffu, my_pix, r);
      if (pix < 0) continue;
      const uint4 sv = *reinterpret_cast<const uint4*>(stage + r * 256 + ((q ^ (r & 7)) << 4));
      float o[4] = {fmaf(P.alpha, __uint_as_float(sv.x), bs[0]), fmaf(P.alpha, __uint_as_float(sv.y), bs[1]),
                    fmaf(P.alpha, __uint_as_float(sv.z), bs[2]), fmaf(P.alpha, __uint_as_float(sv.w), bs[3])};
      const size_t off = (size_t)pix * 128 + hb * 64 + q * 4;
      if (P.res_f32) {
        const float4 rr = *reinterpret_cast<const float4*>(P.res_f32 + off);
        o[0] = fmaf(P.beta, rr.x, o[0]); o[1] = fmaf(P.beta, rr.y, o[1]);
        o[2] = fmaf(P.beta, rr.z, o[2]); o[3] = fmaf(P.beta, rr.w, o[3]);
      }
      if (P.relu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
      }
      if (P.out_f32) *reinterpret_cast<float4*>(P.out_f32 + off) = make_float4(o[0], o[1], o[2], o[3]);
      if (P.out_tf32)
        *reinterpret_cast<float4*>(P.out_tf32 + off) =
            make_float4(round_tf32(o[0]), round_tf32(o[1]), round_tf32(o[2]), round_tf32(o[3]));
    }
    __syncwarp();  // the next half overwrites the staging rows
  }
}

template <int EPI, bool CS = false>
__device__ __forceinline__ void epilogue_acc(const ConvKernelParams& P, const TileCoord& c, uint32_t t_acc,
                                             int f_base, uint8_t* stage_buf, int ew, const float* s_bias, int lane,
                                             bool live, float (&cs)[8]) {
  static_assert(!CS || (EPI >= 1 && EPI <= 3), "column sums ride the residual / mask epilogues");
  // per-warp staging tile: 32 rows x 256 B (512 B in the sub-pixel variant, whose channel table follows the tiles)
  uint8_t* stage = stage_buf + ew * (32 * (EPI == 4 ? kShuffleStageRow : 256));
  if constexpr (EPI < 0) epilogue_staged_acc_generic(P, c, t_acc, f_base, stage, s_bias, lane, live);
  else if constexpr (EPI == 0) epilogue_staged_acc_plain(P, c, t_acc, f_base, stage, s_bias, lane, live);
  else if constexpr (EPI == 4)
    epilogue_shuffle_acc(P, c, t_acc, f_base, stage, reinterpret_cast<const int*>(stage_buf + 4 * 32 * kShuffleStageRow),
                         s_bias, lane, live);
  else if constexpr (EPI == 5) epilogue_staged_acc_tf32(P, c, t_acc, f_base, stage, s_bias, lane, live);
  else epilogue_staged_acc<EPI, CS>(P, c, t_acc, f_base, stage, s_bias, lane, live, cs);
}

// end of an epilogue warp's life: lanes l and l + 16 hold the same eight channels
__device__ __forceinline__ void colsum_flush(const ConvKernelParams& P, float (&cs)[8], int lane) {
#pragma unroll
  for (int e = 0; e < 8; ++e) cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
  if (lane < 16) {
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(P.colsum + lane * 8 + e, P.colsum_scale * cs[e]);
  }
}

}  // namespace

template <int N_, int AMODE, int NACC, int NBUF, int EPI, bool TF32, bool CS>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmW0,
               const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW1,
               const ConvKernelParams P) {
  constexpr int T = NACC * 128;
  constexpr int WTAP = N_ * kRowBytes;       // one tap x one K chunk (32 bf16 / 16 tf32 channels) x N couts
  constexpr int WSTAGE = kTapsPerStage * WTAP;  // a weight stage carries two consecutive taps
  constexpr int NCH = PrecCfg<TF32>::kChunks, CHE = PrecCfg<TF32>::kElems;
  static_assert(!TF32 || AMODE == kAModeSwizzle64, "tf32 operands use the swizzled strip");
  static_assert(!TF32 || N_ != 128 || EPI == 5, "tf32 128-wide launches use the fp32 epilogue");
  constexpr uint32_t TM_COLS_RAW = NACC * NBUF * N_;
  constexpr uint32_t TM_COLS = TM_COLS_RAW <= 32    ? 32
                               : TM_COLS_RAW <= 64  ? 64
                               : TM_COLS_RAW <= 128 ? 128
                               : TM_COLS_RAW <= 256 ? 256
                                                    : 512;
  static_assert(TM_COLS_RAW <= 512, "TMEM overflow");
  constexpr uint32_t IDESC = umma_idesc(TF32 ? 2u : 1u /*tf32 : bf16*/, 128u, (uint32_t)N_);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared space
  uint8_t* a_buf = smem;
  uint8_t* w_buf = smem + 2 * P.a_bytes;
  constexpr int STAGE_BYTES = N_ != 128 ? 0 : EPI == 4 ? kShuffleStageBytes : 4 * 32 * 256;  // epilogue staging: 32 rows x 256 B per warp
  uint8_t* stage_buf = w_buf + P.num_wstages * WSTAGE;
  ConvBarriers* bars = reinterpret_cast<ConvBarriers*>(stage_buf + STAGE_BYTES);
  float* s_bias = reinterpret_cast<float*>(bars + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NS = P.num_wstages;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&bars->w_full[i], 1);
      mbar_init(&bars->w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->a_full[i], 1);
      mbar_init(&bars->a_empty[i], 1);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmW0);
    if (P.nsrc > 1) prefetch_tmap(&tmW1);
  }
  if (warp == 3 && lane == 0) {
    prefetch_tmap(&tmA0);
    if (P.nsrc > 1) prefetch_tmap(&tmA1);
  }
  if (warp == 2) {
    tmem_alloc(&bars->tmem_base, TM_COLS);
    tmem_relinquish();
  }
  griddep_wait();   // everything above overlapped the predecessor's tail; every global access comes after this
  if (threadIdx.x >= 128) {
    for (int i = threadIdx.x - 128; i < N_; i += 128) s_bias[i] = (i < P.cout && P.bias) ? P.bias[i] : 0.f;
    if constexpr (EPI == 4)
      shuffle_table_init(P, reinterpret_cast<int*>(stage_buf + 4 * 32 * kShuffleStageRow), threadIdx.x - 128, 128);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  griddep_launch_dependents();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------ weight-stage TMA producer
    if (lane == 0) {
      Ring wr{0u, 0u, (uint32_t)NS};
      for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
        for (int s = 0; s < P.nsrc; ++s) {
          const CUtensorMap* tmW = s == 0 ? &tmW0 : &tmW1;
          const int ntaps = P.ksize[s] * P.ksize[s];
          for (int ch = 0; ch < P.nch[s]; ++ch) {
            for (int tap = 0; tap < ntaps; tap += kTapsPerStage) {
              // one box = 2 taps (an odd last tap drags in the next 128 rows, unused; OOB rows are zero-filled)
              const uint32_t slot = wr.slot, ph = wr.phase;
              mbar_wait(&bars->w_empty[slot], ph ^ 1);
              mbar_expect_tx(&bars->w_full[slot], WSTAGE);
              tma_load_2d(w_buf + slot * WSTAGE, tmW, &bars->w_full[slot], 0, (ch * ntaps + tap) * N_);
              wr.advance();
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------ activation strip TMA producer
    if (lane == 0) {
      uint32_t ac = 0;
      const uint32_t strip_bytes = (uint32_t)P.NR * P.PWs * kRowBytes;
      for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile<T>(P, t);
        for (int s = 0; s < P.nsrc; ++s) {
          const CUtensorMap* tmA = s == 0 ? &tmA0 : &tmA1;
          for (int ch = 0; ch < P.nch[s]; ++ch) {
            const uint32_t slot = ac & 1, ph = (ac >> 1) & 1;
            mbar_wait(&bars->a_empty[slot], ph ^ 1);
            mbar_expect_tx(&bars->a_full[slot], strip_bytes);
            if constexpr (AMODE == kAModeSwizzle64) {
              tma_load_4d(a_buf + slot * P.a_bytes, tmA, &bars->a_full[slot], ch * CHE,
                          c.seg_x0 - P.p, c.r_lo - P.p, c.n);
            } else {
              tma_load_5d(a_buf + slot * P.a_bytes, tmA, &bars->a_full[slot], 0, c.seg_x0 - P.p,
                          c.r_lo - P.p, ch * (kChunk / 8), c.n);
            }
            ++ac;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    // The loop runs warp-convergent (all 32 lanes wait on the barriers and carry the same descriptor words,
    // so ptxas keeps them in uniform registers); only the elected lane issues tcgen05.mma / commit.
    // Descriptors are {lo, hi} 32-bit words: hi is constant, lo = (smem address >> 4) | LBO field, so a tap
    // shift, an accumulator step (128 rows) and a K step are plain 32-bit adds.
    const bool leader = elect_one();
    uint32_t ac = 0, it = 0;
    Ring wr{0u, 0u, (uint32_t)NS};
    const uint32_t lbo16 = (uint32_t)P.NR * P.PWs;  // interleave mode: bytes between K core matrices, >> 4
    constexpr uint32_t kHiSw64 = (512u >> 4) | (1u << 14) | ((uint32_t)SR_LAYOUT_SW64 << 29);
    constexpr uint32_t kHiNone = (128u >> 4) | (1u << 14) | ((uint32_t)SR_LAYOUT_NONE << 29);
    constexpr uint32_t a_hi = AMODE == kAModeSwizzle64 ? kHiSw64 : kHiNone;
    const uint32_t a_lo_fields = AMODE == kAModeSwizzle64 ? (1u << 16) : (lbo16 << 16);
    const uint32_t a_pix = AMODE == kAModeSwizzle64 ? 4u : 1u;            // descriptor units per pixel row
    const uint32_t a_k16 = AMODE == kAModeSwizzle64 ? 2u : 2u * lbo16;    // descriptor units per K=16 step
    const uint32_t a_buf_lo = (smem_u32(a_buf) >> 4) | a_lo_fields;
    const uint32_t a_slot_step = (uint32_t)P.a_bytes >> 4;
    const uint32_t w_buf_lo = (smem_u32(w_buf) >> 4) | (1u << 16);
    for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
      const TileCoord c = decode_tile<T>(P, t);
      const uint32_t buf = it % NBUF, bph = (it / NBUF) & 1;
      mbar_wait(&bars->tmem_empty[buf], bph ^ 1);
      tc_fence_after();
      const uint32_t d_base = tmem_base + buf * (NACC * N_);
      uint32_t acc_flag = 0;  // 0 only for the first K step of the tile
      for (int s = 0; s < P.nsrc; ++s) {
        const int k = P.ksize[s];
        const int pk = (k - 1) / 2;
        const int ntaps = k * k;
        const int row_wrap = (P.PWs - (k - 1)) * (int)a_pix;  // (ky, k-1) -> (ky+1, 0)
        for (int ch = 0; ch < P.nch[s]; ++ch) {
          const uint32_t aslot = ac & 1, aph = (ac >> 1) & 1;
          mbar_wait(&bars->a_full[aslot], aph);
          tc_fence_after();
          uint32_t a_lo = a_buf_lo + aslot * a_slot_step +
                          (uint32_t)(c.off0 - pk * P.PWs - pk) * a_pix;  // tap (0,0)
          int kx = 0;
          for (int tap = 0; tap < ntaps; tap += kTapsPerStage) {
            const uint32_t wslot = wr.slot, wph = wr.phase;
            mbar_wait(&bars->w_full[wslot], wph);
            tc_fence_after();
            uint32_t b_lo = w_buf_lo + wslot * (uint32_t)(WSTAGE >> 4);
#pragma unroll
            for (int j = 0; j < kTapsPerStage; ++j) {
              if (leader && tap + j < ntaps) {
#pragma unroll
                for (int acc = 0; acc < NACC; ++acc) {
#pragma unroll
                  for (int k16 = 0; k16 < 2; ++k16) {
                    const uint64_t adesc =
                        ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + acc * 128 * a_pix + k16 * a_k16);
                    const uint64_t bdesc = ((uint64_t)kHiSw64 << 32) | (uint64_t)(b_lo + k16 * 2);
                    if constexpr (TF32) umma_tf32(d_base + acc * N_, adesc, bdesc, IDESC, k16 == 0 ? acc_flag : 1u);
                    else umma_bf16(d_base + acc * N_, adesc, bdesc, IDESC, k16 == 0 ? acc_flag : 1u);
                  }
                }
              }
              acc_flag = 1;
              b_lo += (uint32_t)(WTAP >> 4);
              if (++kx == k) {
                kx = 0;
                a_lo += row_wrap;
              } else {
                a_lo += a_pix;
              }
            }
            if (leader) umma_commit(&bars->w_empty[wslot]);
            wr.advance();
          }
          if (leader) umma_commit(&bars->a_empty[aslot]);
          ++ac;
        }
      }
      if (leader) umma_commit(&bars->tmem_full[buf]);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue: TMEM -> registers -> global
    const int ew = warp - 4;  // == warp % 4: TMEM lane quarter this warp may read
    uint32_t it = 0;
    float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x, ++it) {
      const TileCoord c = decode_tile<T>(P, t);
      const uint32_t buf = it % NBUF, bph = (it / NBUF) & 1;
      mbar_wait(&bars->tmem_full[buf], bph);
      tc_fence_after();
      const uint32_t t_base = tmem_base + buf * (NACC * N_) + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
      for (int acc = 0; acc < NACC; ++acc) {
        if constexpr (N_ == 128) {
          epilogue_acc<EPI, CS>(P, c, t_base + acc * N_, c.f0 + acc * 128 + ew * 32, stage_buf, ew, s_bias, lane, true, cs);
        } else {
          const int f = c.f0 + acc * 128 + ew * 32 + lane;
          const int fr = f / P.PWs;
          const int yy = fr - P.p;
          const int xx = f - fr * P.PWs - P.p;
          const bool valid = (f - (P.p * P.PWs + P.p) < P.f_len) && yy < P.Hc && xx >= 0 &&
                             xx < P.BW && (c.seg_x0 + xx) < P.Wc;
          const size_t pix = ((size_t)c.n * P.H + yy) * P.W + c.seg_x0 + xx;
          // optional scatter: image n lands in slot out_index[n] of a tensor of out_H x out_W images
          size_t opix = pix;
          if (P.out_index && valid)
            opix = ((size_t)P.out_index[c.n] * P.out_H + yy) * P.out_W + c.seg_x0 + xx;
          uint32_t v[16];
          tmem_ld16(t_base + acc * N_, v);
          tmem_ld_wait();
          if (valid) {
            // fused stitch: does this patch own the pixel, and does it fall inside the (cropped) image?
            unsigned char* u8 = nullptr;
            if (P.stitch_u8) {
              const sr_stitch_tile tl = P.stitch_tiles[P.out_index ? P.out_index[c.n] : c.n];
              const int px = c.seg_x0 + xx;
              const int Y = tl.y0 + yy, X = tl.x0 + px;
              if (yy >= tl.oy0 && yy < tl.oy1 && px >= tl.ox0 && px < tl.ox1 && Y >= 0 && Y < tl.img_h && X >= 0 &&
                  X < tl.img_w)
                u8 = P.stitch_u8 + tl.img_offset + ((size_t)Y * tl.img_w + X) * 3;
            }
            for (int j = 0; j < P.cout; ++j) {
              float o = P.alpha * (__uint_as_float(v[j]) + s_bias[j]);
              if (P.res_f32) o = fmaf(P.beta, P.res_f32[pix * P.cout + j], o);
              if (P.relu) o = fmaxf(o, 0.f);
              if (P.out_f32) P.out_f32[opix * P.cout + j] = o;
              if (P.out_bf16) P.out_bf16[opix * P.cout + j] = __float2bfloat16_rn(o);
              // np.clip(x * 255, 0, 255).astype('uint8') (models.py:351, 391): the arithmetic of patch_stitch_kernel
              if (u8 && j < 3) u8[j] = (unsigned char)(int)fminf(fmaxf(__fmul_rn(o, P.stitch_mul), 0.f), 255.f);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);
    }
    if constexpr (CS) colsum_flush(P, cs, lane);
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TM_COLS);
  }
}

// =====================================================================================
// CTA-pair variant (cta_group::2): two CTAs of a cluster process the same tile position of two
// different images with ONE tcgen05.mma of M = 256.  Each CTA stages its own activation strip and
// only HALF of every weight stage (64 of the 128 couts); the tensor cores of both SMs read both
// halves, so per-SM shared-memory operand traffic drops from 8 KB to 6 KB per MMA (128 -> 96 B/clk)
// and L2 -> SM weight traffic halves.  The leader CTA (rank 0) owns the full-barriers and issues;
// tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals to both CTAs.
// =====================================================================================
template <int NACC, int NBUF, int EPI, bool TF32, bool CS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmW0,
                    const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW1,
                    const ConvKernelParams P) {
  constexpr int N_ = 128;
  constexpr int T = NACC * 128;
  constexpr int WTAP = (N_ / 2) * kRowBytes;    // this CTA's half (64 couts) of one tap x one K chunk
  constexpr int NCH = PrecCfg<TF32>::kChunks, CHE = PrecCfg<TF32>::kElems;
  static_assert(!TF32 || EPI == 5, "tf32 launches use the fp32 epilogue");
  constexpr int WSTAGE = kTapsPerStage * WTAP;  // a weight stage carries two consecutive taps
  constexpr uint32_t TM_COLS = 512;
  static_assert(NACC * NBUF * N_ == 512, "pair kernel uses the whole TMEM");
  constexpr uint32_t IDESC = umma_idesc(TF32 ? 2u : 1u /*tf32 : bf16*/, 256u, (uint32_t)N_);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared space
  uint8_t* a_buf = smem;
  uint8_t* w_buf = smem + P.num_abuf * P.a_bytes;
  uint8_t* stage_buf = w_buf + P.num_wstages * WSTAGE;
  ConvBarriers* bars = reinterpret_cast<ConvBarriers*>(stage_buf + (EPI == 4 ? kShuffleStageBytes : 4 * 32 * 256));
  float* s_bias = reinterpret_cast<float*>(bars + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NS = P.num_wstages;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  // The two CTAs of a pair take the same tile index of two different COLUMNS (a column = one column segment of
  // one image): their tiles then start at the same offset inside the strip (off0 depends on the tile index only),
  // which is what lets one MMA descriptor address both CTAs' shared memory.  Pairing columns rather than images
  // also pairs the segments of a single image (NB == 1) and leaves no idle CTA when NB is odd but NB * nseg is even.
  const int tps = P.tiles_per_seg;
  const int ncol = P.NB * P.nseg;
  const int pair_tiles = ((ncol + 1) >> 1) * tps;
  if (threadIdx.x == 0) {
    SR_STAMP(P, 0);
    SR_STAMP_NS(P, 14);
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&bars->w_full[i], 1);
      mbar_init(&bars->w_empty[i], 1);
    }
    for (int i = 0; i < P.num_abuf; ++i) {
      mbar_init(&bars->a_full[i], 1);
      mbar_init(&bars->a_empty[i], 1);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], 8);  // 4 epilogue warps of each CTA arrive on the leader's barrier
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmW0);
    if (P.nsrc > 1) prefetch_tmap(&tmW1);
  }
  if (warp == 3 && lane == 0) {
    prefetch_tmap(&tmA0);
    if (P.nsrc > 1) prefetch_tmap(&tmA1);
  }
  if (warp == 2) {
    tmem_alloc_pair(&bars->tmem_base, TM_COLS);
    tmem_relinquish_pair();
  }
  griddep_wait();   // everything above overlapped the predecessor's tail; every global access comes after this
  if (threadIdx.x >= 128) {
    for (int i = threadIdx.x - 128; i < N_; i += 128) s_bias[i] = (i < P.cout && P.bias) ? P.bias[i] : 0.f;
    if constexpr (EPI == 4)
      shuffle_table_init(P, reinterpret_cast<int*>(stage_buf + 4 * 32 * kShuffleStageRow), threadIdx.x - 128, 128);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  griddep_launch_dependents();
  const uint32_t tmem_base = bars->tmem_base;
  if (threadIdx.x == 0) SR_STAMP(P, 1);

  // tile of this CTA for pair-tile index pt: column 2*j + rank (clamped; a clamped duplicate does not store)
  auto decode = [&](int pt, bool* live) {
    const int j = pt / tps;
    const int ti = pt - j * tps;
    int col = 2 * j + (int)rank;
    *live = col < ncol;
    if (col >= ncol) col = ncol - 1;
    return decode_tile<T>(P, col * tps + ti);   // tile order is (image, segment, tile index): col = n * nseg + seg
  };

  if (warp == 0) {
    // ------------------------------------------------ weight half-stage TMA producer (both CTAs)
    if (lane == 0) {
      Ring wr{0u, 0u, (uint32_t)NS};
      bool w_wrapped = false;
      for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters) {
        for (int s = 0; s < P.nsrc; ++s) {
          const CUtensorMap* tmW = s == 0 ? &tmW0 : &tmW1;
          const int ntaps = P.ksize[s] * P.ksize[s];
          for (int ch = 0; ch < P.nch[s]; ++ch) {
            for (int tap = 0; tap < ntaps; tap += kTapsPerStage) {
              const uint32_t slot = wr.slot, ph = wr.phase;
              mbar_wait(&bars->w_empty[slot], ph ^ 1);
              if (SR_DBG(P, 1) && w_wrapped) {   // timing experiment: the stage keeps whatever it held
                if (is_leader) mbar_arrive(&bars->w_full[slot]);
                wr.advance();
                continue;
              }
              if (wr.slot + 1 == wr.n) w_wrapped = true;
              const int nbox = min(kTapsPerStage, ntaps - tap);  // an odd last tap loads one box only
              if (pt == cluster_id && s == 0 && ch == 0 && tap == 0) SR_STAMP(P, 9);
              if (is_leader) mbar_expect_tx(&bars->w_full[slot], 2 * nbox * WTAP);
              const uint32_t bar = mapa_shared(smem_u32(&bars->w_full[slot]), 0);
#pragma unroll
              for (int j = 0; j < kTapsPerStage; ++j)
                if (j < nbox)
                  tma_load_2d_pair(w_buf + slot * WSTAGE + j * WTAP, tmW, bar, 0,
                                   (ch * ntaps + tap + j) * N_ + (int)rank * (N_ / 2));
              wr.advance();
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------ activation strip TMA producer (both CTAs)
    if (lane == 0) {
      Ring ar{0u, 0u, (uint32_t)P.num_abuf};
      bool a_wrapped = false;
      const uint32_t strip_bytes = (uint32_t)P.NR * P.PWs * kRowBytes;
      for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters) {
        bool live;
        const TileCoord c = decode(pt, &live);
        for (int s = 0; s < P.nsrc; ++s) {
          const CUtensorMap* tmA = s == 0 ? &tmA0 : &tmA1;
          for (int ch = 0; ch < P.nch[s]; ++ch) {
            const uint32_t slot = ar.slot, ph = ar.phase;
            mbar_wait(&bars->a_empty[slot], ph ^ 1);
            if (SR_DBG(P, 2) && a_wrapped) {     // timing experiment: the strip buffer keeps whatever it held
              if (is_leader) mbar_arrive(&bars->a_full[slot]);
              ar.advance();
              continue;
            }
            if (ar.slot + 1 == ar.n) a_wrapped = true;
            if (pt == cluster_id && s == 0 && ch == 0) SR_STAMP(P, 10);
            if (is_leader) mbar_expect_tx(&bars->a_full[slot], 2 * strip_bytes);
            tma_load_4d_pair(a_buf + slot * P.a_bytes, tmA, mapa_shared(smem_u32(&bars->a_full[slot]), 0),
                             ch * CHE, c.seg_x0 - P.p, c.r_lo - P.p, c.n);
            ar.advance();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA only, warp-convergent)
    if (is_leader) {
      const bool leader = elect_one();
      uint32_t it = 0;
      Ring wr{0u, 0u, (uint32_t)NS};
      Ring ar{0u, 0u, (uint32_t)P.num_abuf};
      constexpr uint32_t kHi = (512u >> 4) | (1u << 14) | ((uint32_t)SR_LAYOUT_SW64 << 29);
      const uint32_t a_buf_lo = (smem_u32(a_buf) >> 4) | (1u << 16);
      const uint32_t a_slot_step = (uint32_t)P.a_bytes >> 4;
      const uint32_t w_buf_lo = (smem_u32(w_buf) >> 4) | (1u << 16);
      for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters, ++it) {
        bool live;
        const TileCoord c = decode(pt, &live);  // off0 is identical for both CTAs of the pair
        const uint32_t buf = it % NBUF, bph = (it / NBUF) & 1;
        mbar_wait(&bars->tmem_empty[buf], bph ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + buf * (NACC * N_);
        uint32_t acc_flag = 0;
        for (int s = 0; s < P.nsrc; ++s) {
          const int k = P.ksize[s];
          const int pk = (k - 1) / 2;
          const int ntaps = k * k;
          const int row_wrap = (P.PWs - (k - 1)) * 4;
          for (int ch = 0; ch < P.nch[s]; ++ch) {
            const uint32_t aslot = ar.slot, aph = ar.phase;
            mbar_wait(&bars->a_full[aslot], aph);
            tc_fence_after();
            if (it == 0 && s == 0 && leader) SR_STAMP(P, ch == 0 ? 2 : ch == P.nch[s] - 1 ? 12 : 13);
            uint32_t a_lo = a_buf_lo + aslot * a_slot_step + (uint32_t)(c.off0 - pk * P.PWs - pk) * 4u;
            int kx = 0;
            for (int tap = 0; tap < ntaps; tap += kTapsPerStage) {
              const uint32_t wslot = wr.slot, wph = wr.phase;
              mbar_wait(&bars->w_full[wslot], wph);
              tc_fence_after();
              if (it == 0 && s == 0 && ch == 0 && tap == 0 && leader) SR_STAMP(P, 3);
              uint32_t b_lo = w_buf_lo + wslot * (uint32_t)(WSTAGE >> 4);
#pragma unroll
              for (int j = 0; j < kTapsPerStage; ++j) {
                if (leader && tap + j < ntaps) {
#pragma unroll
                  for (int acc = 0; acc < NACC; ++acc) {
#pragma unroll
                    for (int k16 = 0; k16 < 2; ++k16) {
                      const uint64_t adesc = ((uint64_t)kHi << 32) | (uint64_t)(a_lo + acc * 512 + k16 * 2);
                      const uint64_t bdesc = ((uint64_t)kHi << 32) | (uint64_t)(b_lo + k16 * 2);
                      if constexpr (TF32) umma_tf32_pair(d_base + acc * N_, adesc, bdesc, IDESC, k16 == 0 ? acc_flag : 1u);
                      else umma_bf16_pair(d_base + acc * N_, adesc, bdesc, IDESC, k16 == 0 ? acc_flag : 1u);
                    }
                  }
                }
                acc_flag = 1;
                b_lo += (uint32_t)(WTAP >> 4);
                if (++kx == k) {
                  kx = 0;
                  a_lo += row_wrap;
                } else {
                  a_lo += 4;
                }
              }
              if (leader) umma_commit_pair(&bars->w_empty[wslot]);
              wr.advance();
            }
            if (leader) umma_commit_pair(&bars->a_empty[aslot]);
            ar.advance();
          }
        }
        if (leader) umma_commit_pair(&bars->tmem_full[buf]);
        if (it == 0 && leader) SR_STAMP(P, 4);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue (both CTAs, own TMEM lanes = own tile)
    const int ew = warp - 4;
    uint32_t it = 0;
    float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int pt = cluster_id; pt < pair_tiles; pt += num_clusters, ++it) {
      bool live;
      const TileCoord c = decode(pt, &live);
      const uint32_t buf = it % NBUF, bph = (it / NBUF) & 1;
      mbar_wait(&bars->tmem_full[buf], bph);
      tc_fence_after();
      if (it == 0 && ew == 0 && lane == 0) SR_STAMP(P, 5);
      const uint32_t t_base = tmem_base + buf * (NACC * N_) + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
      for (int acc = 0; acc < NACC; ++acc)
        epilogue_acc<EPI, CS>(P, c, t_base + acc * N_, c.f0 + acc * 128 + ew * 32, stage_buf, ew, s_bias, lane, live, cs);
      tc_fence_before();
      __syncwarp();
      if (it == 0 && ew == 0 && lane == 0) SR_STAMP(P, 6);
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&bars->tmem_empty[buf]), 0));
    }
    if constexpr (CS) colsum_flush(P, cs, lane);
  }

  __syncwarp();  // lanes of the single-lane roles reconverge before the aligned cluster barrier
  if (threadIdx.x == 0) SR_STAMP(P, 7);
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TM_COLS);
  }
  if (threadIdx.x == 0) {
    SR_STAMP(P, 8);
    SR_STAMP_NS(P, 15);
  }
}


// =====================================================================================
// Chain kernel: a whole SEQUENCE of convolutions in one persistent launch (small inputs).
//
// A single 128x128 patch (BASELINE config 1) gives every CTA one 128-position tile per layer: a per-layer launch
// then spends more time on launch gap, barrier / TMEM setup, first weight fetch and teardown than on its MMAs
// (tools/probe_timeline.py: 5.7 us of MMA in an 18.6 us k3 launch).  Here the CTA pairs stay resident and walk a list
// of PHASES; a phase is one or two independent convolutions (the two branch heads of a 5/3 block read the same
// tensor), phases are separated by a grid-wide barrier (every output of phase i is in L2 before a strip of phase
// i+1 is requested).  What survives per layer is the barrier (~1.5 us) and the first strip's L2 latency: weights of
// the next phase stream into the ring while the current one computes, TMEM / mbarriers are set up once, and the
// epilogue of the first head tile overlaps the MMAs of the second.
// Same tiles, same tap / chunk / source order per output position as the per-layer kernels: bit-identical results.
// =====================================================================================
struct alignas(128) ChainConv {
  CUtensorMap tmA[2], tmW[2];
  ConvKernelParams P;
  int phase;        // convs of one phase are independent of each other; phases run in order
  int pair_tiles;   // ((NB * nseg + 1) / 2) * tiles_per_seg
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Spin until *counter >= target.  The poll is one acquire load (an L2 round trip); inline and bounded by the SM clock
// like mbar_wait (a protocol bug traps instead of hanging the GPU; no call inside the role loops).
__device__ __forceinline__ void chain_wait_counter(const unsigned* counter, unsigned target) {
  if (ld_acquire_gpu(counter) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(counter) < target) {
    if (clock64() - t0 > SR_MBAR_TIMEOUT_CLK) __trap();
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv_tc_chain_kernel(const ChainConv* __restrict__ convs, int n_convs, int n_phases, unsigned* sync_counter,
                     int a_slot_bytes, int num_abuf, int num_wstages, unsigned long long* tl) {
  constexpr int N_ = 128, NACC = 1, NBUF = 4, T = 128;
  constexpr int WTAP = (N_ / 2) * kRowBytes;
  constexpr int NCH = kNumChunks, CHE = kChunk;
  constexpr int WSTAGE = kTapsPerStage * WTAP;
  constexpr uint32_t TM_COLS = 512;
  constexpr uint32_t IDESC = umma_idesc(1u, 256u, (uint32_t)N_);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_buf = smem;
  uint8_t* w_buf = smem + num_abuf * a_slot_bytes;
  uint8_t* stage_buf = w_buf + num_wstages * WSTAGE;
  ConvBarriers* bars = reinterpret_cast<ConvBarriers*>(stage_buf + 4 * 32 * 256);
  float* s_bias = reinterpret_cast<float*>(bars + 1);   // [4 epilogue warps][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int NS = num_wstages;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&bars->w_full[i], 1);
      mbar_init(&bars->w_empty[i], 1);
    }
    for (int i = 0; i < num_abuf; ++i) {
      mbar_init(&bars->a_full[i], 1);
      mbar_init(&bars->a_empty[i], 1);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(&bars->tmem_full[i], 1);
      mbar_init(&bars->tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(&bars->tmem_base, TM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // development build (sr_dev_set_timeline): globaltimer stamps, 8 per (CTA, phase) -- 0 barrier passed, 1 / 2 first /
  // last strip of the phase's first tile ready, 3 MMAs committed, 4 / 5 epilogue sees the accumulator / is done,
  // 6 fenced, about to arrive (tools/probe_chain.py)
#ifdef SR_DEV_SWITCHES
#define CH_STAMP(ph_, idx_) do { if (tl) tl[((size_t)blockIdx.x * n_phases + (ph_)) * 8 + (idx_)] = global_timer_ns(); } while (0)
#else
#define CH_STAMP(ph_, idx_) do { (void)tl; } while (0)
#endif

  // first pair tile of conv `pt_total` tiles this cluster owns, when the phase's tiles before it number g0
  auto first_tile = [&](int g0) {
    int r = (cluster_id - g0) % num_clusters;
    return r < 0 ? r + num_clusters : r;
  };
  auto decode = [&](const ConvKernelParams& P, int pt, bool* live) {
    const int tps = P.tiles_per_seg, ncol = P.NB * P.nseg;
    const int j = pt / tps;
    const int ti = pt - j * tps;
    int col = 2 * j + (int)rank;
    *live = col < ncol;
    if (col >= ncol) col = ncol - 1;
    return decode_tile<T>(P, col * tps + ti);
  };

  if (warp == 0) {
    // ------------------------------------------------ weight half-stage TMA producer: free-running over all phases
    if (lane == 0) {
      Ring wr{0u, 0u, (uint32_t)NS};
      int g0 = 0, ph = 0;
      for (int ci = 0; ci < n_convs; ++ci) {
        const ChainConv& cv = convs[ci];
        if (cv.phase != ph) { ph = cv.phase; g0 = 0; }
        const ConvKernelParams& P = cv.P;
        for (int pt = first_tile(g0); pt < cv.pair_tiles; pt += num_clusters) {
          for (int s = 0; s < P.nsrc; ++s) {
            const CUtensorMap* tmW = &cv.tmW[s];
            const int ntaps = P.ksize[s] * P.ksize[s];
            for (int ch = 0; ch < NCH; ++ch) {
              for (int tap = 0; tap < ntaps; tap += kTapsPerStage) {
                const uint32_t slot = wr.slot, wph = wr.phase;
                mbar_wait(&bars->w_empty[slot], wph ^ 1);
                const int nbox = min(kTapsPerStage, ntaps - tap);
                if (is_leader) mbar_expect_tx(&bars->w_full[slot], 2 * nbox * WTAP);
                const uint32_t bar = mapa_shared(smem_u32(&bars->w_full[slot]), 0);
#pragma unroll
                for (int j = 0; j < kTapsPerStage; ++j)
                  if (j < nbox)
                    tma_load_2d_pair(w_buf + slot * WSTAGE + j * WTAP, tmW, bar, 0,
                                     (ch * ntaps + tap + j) * N_ + (int)rank * (N_ / 2));
                wr.advance();
              }
            }
          }
        }
        g0 += cv.pair_tiles;
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------ activation strip TMA producer; waits for the grid barrier
    if (lane == 0) {
      Ring ar{0u, 0u, (uint32_t)num_abuf};
      int g0 = 0, ph = 0;
      for (int ci = 0; ci < n_convs; ++ci) {
        const ChainConv& cv = convs[ci];
        if (cv.phase != ph) {
          ph = cv.phase;
          g0 = 0;
          // every CTA of the grid has stored (and fenced) its outputs of phase ph - 1
          chain_wait_counter(sync_counter, (unsigned)ph * gridDim.x);
          CH_STAMP(ph, 0);
          fence_proxy_async_all();   // generic-proxy stores of other SMs -> this thread's async-proxy (TMA) reads
        }
        const ConvKernelParams& P = cv.P;
        const uint32_t strip_bytes = (uint32_t)P.NR * P.PWs * kRowBytes;
        for (int pt = first_tile(g0); pt < cv.pair_tiles; pt += num_clusters) {
          bool live;
          const TileCoord c = decode(P, pt, &live);
          for (int s = 0; s < P.nsrc; ++s) {
            const CUtensorMap* tmA = &cv.tmA[s];
            for (int ch = 0; ch < NCH; ++ch) {
              const uint32_t slot = ar.slot, aph = ar.phase;
              mbar_wait(&bars->a_empty[slot], aph ^ 1);
              if (is_leader) mbar_expect_tx(&bars->a_full[slot], 2 * strip_bytes);
              tma_load_4d_pair(a_buf + slot * a_slot_bytes, tmA, mapa_shared(smem_u32(&bars->a_full[slot]), 0),
                               ch * CHE, c.seg_x0 - P.p, c.r_lo - P.p, c.n);
              ar.advance();
            }
          }
        }
        g0 += cv.pair_tiles;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (leader CTA only, warp-convergent)
    if (is_leader) {
      const bool leader = elect_one();
      uint32_t it = 0;
      Ring wr{0u, 0u, (uint32_t)NS};
      Ring ar{0u, 0u, (uint32_t)num_abuf};
      constexpr uint32_t kHi = (512u >> 4) | (1u << 14) | ((uint32_t)SR_LAYOUT_SW64 << 29);
      const uint32_t a_buf_lo = (smem_u32(a_buf) >> 4) | (1u << 16);
      const uint32_t a_slot_step = (uint32_t)a_slot_bytes >> 4;
      const uint32_t w_buf_lo = (smem_u32(w_buf) >> 4) | (1u << 16);
      int g0 = 0, ph = 0;
      for (int ci = 0; ci < n_convs; ++ci) {
        const ChainConv& cv = convs[ci];
        if (cv.phase != ph) { ph = cv.phase; g0 = 0; }
        const ConvKernelParams& P = cv.P;
        for (int pt = first_tile(g0); pt < cv.pair_tiles; pt += num_clusters, ++it) {
          bool live;
          const TileCoord c = decode(P, pt, &live);
          const uint32_t buf = it % NBUF, bph = (it / NBUF) & 1;
          mbar_wait(&bars->tmem_empty[buf], bph ^ 1);
          tc_fence_after();
          const uint32_t d_base = tmem_base + buf * (NACC * N_);
          uint32_t acc_flag = 0;
          for (int s = 0; s < P.nsrc; ++s) {
            const int k = P.ksize[s];
            const int pk = (k - 1) / 2;
            const int ntaps = k * k;
            const int row_wrap = (P.PWs - (k - 1)) * 4;
            for (int ch = 0; ch < NCH; ++ch) {
              const uint32_t aslot = ar.slot, aph = ar.phase;
              mbar_wait(&bars->a_full[aslot], aph);
              tc_fence_after();
              if (leader && s == 0 && ch == 0) CH_STAMP(cv.phase, 1);
              if (leader && s == P.nsrc - 1 && ch == NCH - 1) CH_STAMP(cv.phase, 2);
              uint32_t a_lo = a_buf_lo + aslot * a_slot_step + (uint32_t)(c.off0 - pk * P.PWs - pk) * 4u;
              int kx = 0;
              for (int tap = 0; tap < ntaps; tap += kTapsPerStage) {
                const uint32_t wslot = wr.slot, wph = wr.phase;
                mbar_wait(&bars->w_full[wslot], wph);
                tc_fence_after();
                uint32_t b_lo = w_buf_lo + wslot * (uint32_t)(WSTAGE >> 4);
#pragma unroll
                for (int j = 0; j < kTapsPerStage; ++j) {
                  if (leader && tap + j < ntaps) {
#pragma unroll
                    for (int k16 = 0; k16 < 2; ++k16) {
                      const uint64_t adesc = ((uint64_t)kHi << 32) | (uint64_t)(a_lo + k16 * 2);
                      const uint64_t bdesc = ((uint64_t)kHi << 32) | (uint64_t)(b_lo + k16 * 2);
                      umma_bf16_pair(d_base, adesc, bdesc, IDESC, k16 == 0 ? acc_flag : 1u);
                    }
                  }
                  acc_flag = 1;
                  b_lo += (uint32_t)(WTAP >> 4);
                  if (++kx == k) {
                    kx = 0;
                    a_lo += row_wrap;
                  } else {
                    a_lo += 4;
                  }
                }
                if (leader) umma_commit_pair(&bars->w_empty[wslot]);
                wr.advance();
              }
              if (leader) umma_commit_pair(&bars->a_empty[aslot]);
              ar.advance();
            }
          }
          if (leader) umma_commit_pair(&bars->tmem_full[buf]);
          if (leader) CH_STAMP(cv.phase, 3);
        }
        g0 += cv.pair_tiles;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ epilogue (both CTAs); arrives at the grid barrier per phase
    const int ew = warp - 4;
    uint32_t it = 0;
    float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float* my_bias = s_bias + ew * N_;
    int g0 = 0, ph = 0;
    auto end_phase = [&]() {
      // this CTA's outputs of the phase are stored: make them visible GPU-wide (also to the async proxy of the
      // SMs that will TMA-load them), then one arrival per CTA
      __threadfence();
      fence_proxy_async_all();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (ew == 0 && lane == 0) CH_STAMP(ph, 6);
      if (ew == 0 && lane == 0) {
        // A CTA without tiles in this phase must not run ahead: it may only arrive for phase ph once every CTA has
        // arrived for phase ph - 1 (for a CTA with tiles that already holds -- its strips waited for it).  Then
        // "counter >= (ph + 1) * gridDim.x" really means that every CTA has finished phase ph.
        chain_wait_counter(sync_counter, (unsigned)ph * gridDim.x);
        red_release_gpu_add(sync_counter, 1u);
      }
    };
    for (int ci = 0; ci < n_convs; ++ci) {
      const ChainConv& cv = convs[ci];
      if (cv.phase != ph) {
        end_phase();
        ph = cv.phase;
        g0 = 0;
      }
      const ConvKernelParams& P = cv.P;
      bool bias_loaded = false;
      for (int pt = first_tile(g0); pt < cv.pair_tiles; pt += num_clusters, ++it) {
        if (!bias_loaded) {
          __syncwarp();
          for (int i = lane; i < N_; i += 32) my_bias[i] = (i < P.cout && P.bias) ? P.bias[i] : 0.f;
          __syncwarp();
          bias_loaded = true;
        }
        bool live;
        const TileCoord c = decode(P, pt, &live);
        const uint32_t buf = it % NBUF, bph = (it / NBUF) & 1;
        mbar_wait(&bars->tmem_full[buf], bph);
        tc_fence_after();
        if (ew == 0 && lane == 0) CH_STAMP(cv.phase, 4);
        const uint32_t t_base = tmem_base + buf * (NACC * N_) + ((uint32_t)(ew * 32) << 16);
        uint8_t* stage = stage_buf + ew * (32 * 256);
        if (P.res_f32) epilogue_staged_acc<1, false>(P, c, t_base, c.f0 + ew * 32, stage, my_bias, lane, live, cs);
        else epilogue_staged_acc_plain(P, c, t_base, c.f0 + ew * 32, stage, my_bias, lane, live);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&bars->tmem_empty[buf]), 0));
        if (ew == 0 && lane == 0) CH_STAMP(cv.phase, 5);
      }
      g0 += cv.pair_tiles;
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TM_COLS);
  }
#undef CH_STAMP
}

// =====================================================================================
// Host side: plan
// =====================================================================================

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

struct ConvPlan {
  CUtensorMap tmA[2], tmW[2];
  ConvKernelParams P;
  int n_pad;     // 128 or 16
  int amode;
  int nacc;
  int pair;     // 1: CTA-pair (cta_group::2) kernel
  int tf32;     // 1: fp32 tensors, tf32 MMAs
  int grid;
  size_t smem_bytes;
  double flops;  // algorithmic FLOPs (2*MAC) of one run
};

static int make_a_map(CUtensorMap* tm, const void* ptr, int NB, int H, int W, int PWs, int NR,
                      int amode, int tf32) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(SR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  CUresult r;
  if (tf32) {
    // fp32 NHWC: a K chunk is 16 channels = the same 64-byte pixel rows as 32 bf16 channels
    cuuint64_t dims[4] = {(cuuint64_t)kCin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint64_t strides[3] = {(cuuint64_t)kCin * 4, (cuuint64_t)W * kCin * 4,
                             (cuuint64_t)H * W * kCin * 4};
    cuuint32_t box[4] = {(cuuint32_t)PrecCfg<true>::kElems, (cuuint32_t)PWs, (cuuint32_t)NR, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (amode == kAModeSwizzle64) {
    cuuint64_t dims[4] = {(cuuint64_t)kCin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NB};
    cuuint64_t strides[3] = {(cuuint64_t)kCin * 2, (cuuint64_t)W * kCin * 2,
                             (cuuint64_t)H * W * kCin * 2};
    cuuint32_t box[4] = {(cuuint32_t)kChunk, (cuuint32_t)PWs, (cuuint32_t)NR, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)kCin / 8, (cuuint64_t)NB};
    cuuint64_t strides[4] = {(cuuint64_t)kCin * 2, (cuuint64_t)W * kCin * 2, 16,
                             (cuuint64_t)H * W * kCin * 2};
    cuuint32_t box[5] = {8, (cuuint32_t)PWs, (cuuint32_t)NR, (cuuint32_t)kChunk / 8, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(A) failed: %d (W=%d H=%d PWs=%d NR=%d)",
             (int)r, W, H, PWs, NR);
    return set_error(SR_ERR_CUDA, msg);
  }
  return SR_OK;
}

static int make_w_map(CUtensorMap* tm, const void* ptr, int nstages, int n_pad, int box_rows, int tf32) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(SR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int elems = tf32 ? PrecCfg<true>::kElems : kChunk;  // 64-byte rows either way
  cuuint64_t dims[2] = {(cuuint64_t)elems, (cuuint64_t)nstages * n_pad};
  cuuint64_t strides[1] = {(cuuint64_t)kRowBytes};
  cuuint32_t box[2] = {(cuuint32_t)elems, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(tm, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[128];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(W) failed: %d", (int)r);
    return set_error(SR_ERR_CUDA, msg);
  }
  return SR_OK;
}

static constexpr size_t kSmemBudget = 227 * 1024;

// Choose the column-segment width: maximise useful MMA rows subject to the shared-memory budget.
// max_cols_tiles > 0 (sub-wave launches, every CTA pair gets at most one tile): among the geometries whose
// NB * nseg * tiles_per_seg stays <= max_cols_tiles choose the one with the FEWEST STRIP BYTES per tile instead of the
// best MMA row efficiency -- such a launch is bound by the burst of strip loads after it starts (a 128-position tile of
// a 66-pixel-wide segment fetches 5-6 strip rows for its 2 image rows; narrower segments cut that by 28-40 %), not by
// the discarded halo positions.
static bool choose_geometry(int H, int W, int p, int T, int wstage, int stage_bytes,
                            ConvKernelParams* P, int min_nseg = 1, int nseg_step = 1, int nabuf = 2,
                            int NB = 1, long long max_cols_tiles = 0) {
  double best_eff = -1.0;
  for (int nseg = min_nseg; nseg <= W; nseg += nseg_step) {
    const int BW = (W + nseg - 1) / nseg;
    if ((BW * (nseg - 1)) >= W) continue;  // last segment would be empty
    const int PWs = BW + 2 * p;
    if (PWs > 256) continue;
    const int NR = (2 * p * PWs + 2 * p + PWs + T - 2) / PWs + 1;
    if (NR > 256) continue;
    const size_t a_bytes = ((size_t)NR * PWs * kChunk * 2 + 1023) & ~(size_t)1023;
    const size_t fixed =
        nabuf * a_bytes + 1024 /*align slack*/ + stage_bytes + sizeof(ConvBarriers) + 128 * 4 + 64;
    if (fixed + 3 * (size_t)wstage > kSmemBudget) continue;  // >= 3 two-tap weight stages in flight
    const int f_len = (H - 1) * PWs + BW;
    const int tps = (f_len + T - 1) / T;
    double eff = (double)H * W / ((double)nseg * tps * T);
    if (max_cols_tiles > 0) {
      if ((long long)NB * nseg * tps > max_cols_tiles) continue;
      eff = 1.0 / ((double)NR * PWs);            // fewest strip pixels wins
    }
    // prefer fewer segments on ties (less halo traffic)
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      P->BW = BW;
      P->nseg = nseg;
      P->PWs = PWs;
      P->NR = NR;
      P->a_bytes = (int)a_bytes;
      P->num_abuf = nabuf;
      P->f_len = f_len;
      P->tiles_per_seg = tps;
      int ns = (int)((kSmemBudget - fixed) / wstage);
      P->num_wstages = std::min(ns, kMaxWStages);
    }
    if (BW <= 16) break;
  }
  return best_eff > 0;
}

template <int N_, int AMODE, int NACC, int NBUF, int EPI = -1, bool TF32 = false, bool CS = false>
static int launch_variant(const ConvPlan* pl, cudaStream_t stream) {
  auto kern = conv_tc_kernel<N_, AMODE, NACC, NBUF, EPI, TF32, CS>;
  static unsigned long long attr_done = 0;
  if (int rc = ensure_dynamic_smem(kern, (int)kSmemBudget, &attr_done, "cudaFuncSetAttribute(conv_tc_kernel)")) return rc;
  cudaError_t e = launch_pdl(kern, pl->grid, kConvThreads, pl->smem_bytes, stream, pl->tmA[0], pl->tmW[0], pl->tmA[1],
                             pl->tmW[1], pl->P);
  if (e != cudaSuccess) return set_cuda_error(e, "conv_tc_kernel launch");
  return SR_OK;
}

template <int NACC, int NBUF, int EPI = -1, bool TF32 = false, bool CS = false>
static int launch_pair(const ConvPlan* pl, cudaStream_t stream) {
  auto kern = conv_tc_pair_kernel<NACC, NBUF, EPI, TF32, CS>;
  static unsigned long long attr_done = 0;
  if (int rc = ensure_dynamic_smem(kern, (int)kSmemBudget, &attr_done, "cudaFuncSetAttribute(conv_tc_pair_kernel)")) return rc;
  cudaError_t e = launch_pdl(kern, pl->grid, kConvThreads, pl->smem_bytes, stream, pl->tmA[0], pl->tmW[0], pl->tmA[1],
                             pl->tmW[1], pl->P);
  if (e != cudaSuccess) return set_cuda_error(e, "conv_tc_pair_kernel launch");
  return SR_OK;
}

}  // namespace sr

using namespace sr;

// epilogue specialisation: which single global operand it reads (-1: generic run-time epilogue)
static int epilogue_kind(const ConvKernelParams& P) {
  const int nops = (P.res_f32 ? 1 : 0) + ((P.res_bf16 && !P.res_f32) ? 1 : 0) + (P.relu_mask_bf16 ? 1 : 0);
  const bool plain = nops == 0 && !P.shuffle_r && P.out_bf16 && !P.out_f32 && P.relu != 2;
  return P.shuffle_r ? 4 : plain ? 0 : (nops != 1 || P.relu == 2 || P.mask_slope != 0.f) ? -1 : P.res_f32 ? 1 : P.res_bf16 ? 2 : 3;
}

extern "C" int sr_conv_plan_create(const sr_conv_desc* d, sr_conv_plan** out) {
  if (!d || !out) return set_error(SR_ERR_INVALID, "sr_conv_plan_create: null argument");
  if (d->nsrc < 1 || d->nsrc > 2) return set_error(SR_ERR_INVALID, "nsrc must be 1 or 2");
  if (d->cin != kCin) return set_error(SR_ERR_UNSUPPORTED, "tensor-core conv requires cin == 128");
  if (!(d->cout == 128 || (d->cout >= 1 && d->cout <= 16) || (d->shuffle_r > 0 && d->cout > 16 && d->cout < 128)))
    return set_error(SR_ERR_UNSUPPORTED, "tensor-core conv supports cout == 128 or cout <= 16 (any cout <= 128 with shuffle_r)");
  if (d->shuffle_r > 0 && (d->cout % (d->shuffle_r * d->shuffle_r) != 0 || !d->out_f32 || d->out_bf16 ||
                           d->shuffle_order < 0 || d->shuffle_order > 2))
    return set_error(SR_ERR_INVALID, "shuffle_r needs cout divisible by r*r, an fp32 output only, order 0..2");
  if (d->shuffle_r > 0 && (d->cout <= 16 || d->res_f32 || d->res_bf16 || d->relu_mask_bf16 || d->a_mode == 1))
    return set_error(SR_ERR_UNSUPPORTED, "shuffle_r: 16 < cout <= 128, no residual / mask, a_mode 0");
  if (d->NB < 1 || d->H < 1 || d->W < 1) return set_error(SR_ERR_INVALID, "empty tensor");
  if (d->relu < 0 || d->relu > 2) return set_error(SR_ERR_INVALID, "relu must be 0, 1 (ReLU) or 2 (LeakyReLU)");
  if (d->relu == 2 && (d->cout != 128 || d->shuffle_r > 0))
    return set_error(SR_ERR_UNSUPPORTED, "LeakyReLU epilogue: cout == 128, no shuffle");
  if (d->colsum_f32 && (d->cout != 128 || !d->out_bf16 || d->precision == 1 || d->shuffle_r > 0 || d->a_mode == 1 ||
                        d->relu == 2 || d->nacc != 2 || (d->relu_mask_bf16 && d->relu_mask_slope != 0.f) ||
                        ((d->res_f32 ? 1 : 0) + ((d->res_bf16 && !d->res_f32) ? 1 : 0) + (d->relu_mask_bf16 ? 1 : 0)) != 1))
    return set_error(SR_ERR_UNSUPPORTED, "colsum_f32: bf16 128-channel output with exactly one of residual / relu mask, nacc 2, a_mode 0");
  const int tf32 = d->precision == 1 ? 1 : 0;
  if (d->precision != 0 && d->precision != 1) return set_error(SR_ERR_INVALID, "precision must be 0 (bf16) or 1 (tf32)");
  if (tf32 && (d->out_bf16 || d->res_bf16 || d->relu_mask_bf16 || d->shuffle_r > 0 || d->a_mode == 1 || d->relu == 2))
    return set_error(SR_ERR_UNSUPPORTED, "tf32 precision: fp32 tensors only (out_f32 / out_tf32 / res_f32), a_mode 0, no shuffle / mask / LeakyReLU");
  if (!tf32 && d->out_tf32) return set_error(SR_ERR_INVALID, "out_tf32 needs precision == 1");
  if (tf32 && d->cout <= 16 && d->out_tf32) return set_error(SR_ERR_UNSUPPORTED, "out_tf32 needs cout == 128");
  int p = 0;
  for (int s = 0; s < d->nsrc; ++s) {
    const int k = d->ksize[s];
    if (!(k == 1 || k == 3 || k == 5 || k == 7))
      return set_error(SR_ERR_UNSUPPORTED, "kernel size must be 1, 3, 5 or 7");
    if (!d->in[s] || !d->wpacked[s]) return set_error(SR_ERR_INVALID, "null source pointer");
    p = std::max(p, (k - 1) / 2);
  }
  ConvPlan* pl = new (std::nothrow) ConvPlan();
  if (!pl) return set_error(SR_ERR_NOMEM, "out of host memory");
  memset(pl, 0, sizeof *pl);
  pl->n_pad = (d->cout == 128 || d->shuffle_r > 0) ? 128 : 16;
  pl->amode = d->a_mode == 1 ? kAModeInterleave : kAModeSwizzle64;
  pl->tf32 = tf32;
  pl->nacc = ((d->nacc == 2 || d->shuffle_r > 0 || tf32) && pl->n_pad == 128) ? 2 : 4;
  const int T = pl->nacc * 128;
  // CTA pairs work on two columns (image x column segment).  A single image could be split into two segments for
  // it (the kernel handles that), but measured on BASELINE config 1 the extra halo of the forced split costs what
  // the paired MMAs gain (2.47 vs 2.41 ms), so single images keep the single-CTA kernel.
  // A single WIDE image is different: once the strip pitch limit (256 pixels) splits it into >= 2 segments anyway,
  // the segments pair up at no extra cost (the HR stage of config 1: 512 wide = 4 segments of 128).
  const bool pair_ok = d->pair && pl->n_pad == 128 && pl->amode == kAModeSwizzle64;
  pl->pair = (pair_ok && d->NB >= 2) ? 1 : 0;
  int wstage = kTapsPerStage * (pl->pair ? pl->n_pad / 2 : pl->n_pad) * kRowBytes;
  ConvKernelParams& P = pl->P;
  P.nsrc = d->nsrc;
  P.ksize[0] = d->ksize[0];
  P.ksize[1] = d->nsrc > 1 ? d->ksize[1] : 0;
  P.H = d->H;
  P.W = d->W;
  P.NB = d->NB;
  P.p = p;
  const int stage_bytes = pl->n_pad != 128 ? 0 : d->shuffle_r > 0 ? kShuffleStageBytes : 4 * 32 * 256;
  if ((double)d->NB * d->H * d->W >= 2147483647.0) {
    delete pl;
    return set_error(SR_ERR_UNSUPPORTED, "more than 2^31 pixels per tensor");
  }
  // compute extents: only the top-left comp_h x comp_w corner of every image is produced (the input tensor keeps
  // its real data beyond it, so nothing but the true image border is zero-padded)
  const int Hc = (d->comp_h > 0 && d->comp_h < d->H) ? d->comp_h : d->H;
  const int Wc = (d->comp_w > 0 && d->comp_w < d->W) ? d->comp_w : d->W;
  P.Hc = Hc;
  P.Wc = Wc;
  bool geo = choose_geometry(Hc, Wc, p, T, wstage, stage_bytes, &P);
  if (geo && !pl->pair && pair_ok && d->NB == 1 && P.nseg >= 2) {
    // the single image splits into segments by itself: take the pair kernel on an even number of segments
    ConvKernelParams Q = P;
    const int wstage2 = kTapsPerStage * (pl->n_pad / 2) * kRowBytes;
    if (choose_geometry(Hc, Wc, p, T, wstage2, stage_bytes, &Q, 2, 2)) {   // an even number of segments: no idle CTA
      P = Q;
      pl->pair = 1;
      wstage = wstage2;
    }
  }
  if (!geo) {
    delete pl;
    return set_error(SR_ERR_UNSUPPORTED, "no conv geometry fits shared memory");
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  P.res_f32 = d->res_f32;
  P.res_bf16 = reinterpret_cast<const __nv_bfloat16*>(d->res_bf16);
  P.out_bf16 = reinterpret_cast<__nv_bfloat16*>(d->out_bf16);
  P.out_f32 = d->out_f32;
  P.relu = d->relu;
  P.relu_mask_bf16 = reinterpret_cast<const __nv_bfloat16*>(d->relu_mask_bf16);
  P.mask_slope = d->relu_mask_bf16 ? d->relu_mask_slope : 0.f;
  P.shuffle_r = d->shuffle_r > 0 ? d->shuffle_r : 0;
  // Launches of less than one wave (a single small image: BASELINE config 1's LR stage is 65 tiles of 256 positions
  // on 148 SMs): halve the tile to T = 128 (one accumulator, four TMEM buffers) on CTA pairs -- a single image is
  // split into two column segments for it -- so twice as many SMs share the MMAs while every CTA still fetches only
  // half of each weight stage; the strips of all four K chunks are in flight at once (num_abuf up to 4): a tile's
  // MMA phase is too short to hide a strip load behind it.
  if (pair_ok && !tf32 && pl->nacc == 2 && !d->colsum_f32) {
    const int ek = epilogue_kind(P);
    const int clusters = sms / 2;
    const long long cur_tiles = pl->pair ? (long long)((P.NB * P.nseg + 1) / 2) * P.tiles_per_seg
                                         : (long long)P.NB * P.nseg * P.tiles_per_seg;
    const long long cur_waves = pl->pair ? (cur_tiles + clusters - 1) / clusters : (cur_tiles + sms - 1) / sms;
    if ((ek == 0 || ek == 1) && cur_waves == 1) {
      ConvKernelParams Q = P;
      const int wstage2 = kTapsPerStage * (pl->n_pad / 2) * kRowBytes;
      bool ok = false;
      for (int nab = 4; nab >= 2 && !ok; --nab) {
        // first choice: one tile per CTA pair with the fewest strip bytes (2 * clusters tiles at most)
        ok = (d->NB >= 2) ? choose_geometry(Hc, Wc, p, 128, wstage2, stage_bytes, &Q, 1, 1, nab, d->NB, 2LL * clusters)
                          : choose_geometry(Hc, Wc, p, 128, wstage2, stage_bytes, &Q, 2, 2, nab, d->NB, 2LL * clusters);
        if (!ok)
          ok = (d->NB >= 2) ? choose_geometry(Hc, Wc, p, 128, wstage2, stage_bytes, &Q, 1, 1, nab)
                            : choose_geometry(Hc, Wc, p, 128, wstage2, stage_bytes, &Q, 2, 2, nab);
        if (ok && Q.num_wstages < 6 && nab > 2) ok = false;    // keep a useful weight ring
      }
      if (ok) {
        const long long t128 = (long long)((Q.NB * Q.nseg + 1) / 2) * Q.tiles_per_seg;
        const long long w128 = (t128 + clusters - 1) / clusters;
        if (w128 * 128 < cur_waves * 256) {
          P = Q;
          pl->pair = 1;
          pl->nacc = 1;
          wstage = wstage2;
        }
      }
    }
  }
  P.total_tiles = P.NB * P.nseg * P.tiles_per_seg;
  P.bias = d->bias;
  P.alpha = d->alpha;
  P.beta = d->beta;
  P.relu = d->relu;
  P.neg_slope = d->leaky_slope;
  P.res_f32 = d->res_f32;
  P.res_bf16 = reinterpret_cast<const __nv_bfloat16*>(d->res_bf16);
  P.out_bf16 = reinterpret_cast<__nv_bfloat16*>(d->out_bf16);
  P.out_f32 = d->out_f32;
  P.out_tf32 = d->out_tf32;
  {
    const char* e = dev_getenv("SR100_CONV_DBG");
    P.dbg = e ? atoi(e) : 0;
    P.timeline = dev_timeline();
  }
  P.colsum = d->colsum_f32;
  P.colsum_scale = d->colsum_scale;
  P.cout = d->cout;
  P.relu_mask_bf16 = reinterpret_cast<const __nv_bfloat16*>(d->relu_mask_bf16);
  P.out_index = d->out_index;
  P.stitch_tiles = d->stitch_u8 ? d->stitch_tiles : nullptr;
  P.stitch_u8 = d->stitch_u8;
  P.stitch_mul = d->stitch_mul;
  P.shuffle_r = d->shuffle_r > 0 ? d->shuffle_r : 0;
  P.shuffle_order = d->shuffle_order;
  P.shuffle_C = d->shuffle_r > 0 ? d->cout / (d->shuffle_r * d->shuffle_r) : 0;
  P.out_H = d->out_h;
  P.out_W = d->out_w;
  if (d->out_index && (pl->n_pad == 128 || (d->out_f32 && (d->out_h < d->H || d->out_w < d->W)))) {
    delete pl;
    return set_error(SR_ERR_UNSUPPORTED, "out_index scatter needs cout <= 16 and out_h/out_w >= H/W");
  }
  if (d->stitch_u8 && (pl->n_pad == 128 || d->cout != 3 || !d->stitch_tiles)) {
    delete pl;
    return set_error(SR_ERR_UNSUPPORTED, "fused stitch needs cout == 3 and a stitch_tiles table");
  }
  double macs = 0;
  for (int s = 0; s < 2; ++s) {   // K chunks per source: channels >= cin_valid are zeros the launch never multiplies
    const int che = tf32 ? PrecCfg<true>::kElems : kChunk;
    const int cv = (s < d->nsrc && d->cin_valid[s] > 0) ? d->cin_valid[s] : kCin;
    if (cv > kCin || cv % che != 0) {
      delete pl;
      return set_error(SR_ERR_INVALID, "cin_valid must be a multiple of the K chunk (32 bf16 / 16 tf32 channels), <= 128");
    }
    P.nch[s] = cv / che;
  }
  for (int s = 0; s < d->nsrc; ++s) {
    int rc = make_a_map(&pl->tmA[s], d->in[s], d->NB, d->H, d->W, P.PWs, P.NR, pl->amode, tf32);
    if (rc == SR_OK)
      rc = make_w_map(&pl->tmW[s], d->wpacked[s],
                      (tf32 ? PrecCfg<true>::kChunks : kNumChunks) * d->ksize[s] * d->ksize[s], pl->n_pad,
                      pl->pair ? pl->n_pad / 2 : kTapsPerStage * pl->n_pad, tf32);
    if (rc != SR_OK) {
      delete pl;
      return rc;
    }
    macs += (double)d->NB * Hc * Wc * d->ksize[s] * d->ksize[s] * (d->cin_valid[s] > 0 ? d->cin_valid[s] : kCin) * d->cout;
  }
  if (d->nsrc == 1) {
    pl->tmA[1] = pl->tmA[0];
    pl->tmW[1] = pl->tmW[0];
  }
  pl->flops = 2.0 * macs;
  if (pl->pair) {
    const int pair_tiles = ((P.NB * P.nseg + 1) / 2) * P.tiles_per_seg;
    pl->grid = 2 * std::min(pair_tiles, sms / 2);
  } else {
    pl->grid = std::min(P.total_tiles, sms);
  }
  pl->smem_bytes = 1024 + (size_t)P.num_abuf * P.a_bytes + (size_t)P.num_wstages * wstage + stage_bytes +
                   sizeof(ConvBarriers) + 128 * 4 + 64;
  *out = reinterpret_cast<sr_conv_plan*>(pl);
  return SR_OK;
}

extern "C" int sr_conv_plan_run(sr_conv_plan* plan, void* stream) {
  if (!plan) return set_error(SR_ERR_INVALID, "sr_conv_plan_run: null plan");
  const ConvPlan* pl = reinterpret_cast<const ConvPlan*>(plan);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // epilogue specialisation: which single global operand it reads (-1: generic run-time epilogue)
  const ConvKernelParams& P = pl->P;
  const int epi = epilogue_kind(P);
  if (pl->tf32) {
    if (pl->pair) return launch_pair<2, 2, 5, true>(pl, st);
    if (pl->n_pad == 128) return launch_variant<128, kAModeSwizzle64, 2, 2, 5, true>(pl, st);
    return launch_variant<16, kAModeSwizzle64, 4, 2, -1, true>(pl, st);
  }
  if (P.colsum) {   // training: column sums of the stored gradient ride the residual / mask epilogues
    if (pl->pair) {
      switch (epi) {
        case 1: return launch_pair<2, 2, 1, false, true>(pl, st);
        case 2: return launch_pair<2, 2, 2, false, true>(pl, st);
        default: return launch_pair<2, 2, 3, false, true>(pl, st);
      }
    }
    switch (epi) {
      case 1: return launch_variant<128, kAModeSwizzle64, 2, 2, 1, false, true>(pl, st);
      case 2: return launch_variant<128, kAModeSwizzle64, 2, 2, 2, false, true>(pl, st);
      default: return launch_variant<128, kAModeSwizzle64, 2, 2, 3, false, true>(pl, st);
    }
  }
  if (pl->pair) {
    if (pl->nacc == 4) return launch_pair<4, 1>(pl, st);
    if (pl->nacc == 1) return epi == 0 ? launch_pair<1, 4, 0>(pl, st) : launch_pair<1, 4, 1>(pl, st);
    switch (epi) {
      case 0: return launch_pair<2, 2, 0>(pl, st);
      case 1: return launch_pair<2, 2, 1>(pl, st);
      case 2: return launch_pair<2, 2, 2>(pl, st);
      case 3: return launch_pair<2, 2, 3>(pl, st);
      case 4: return launch_pair<2, 2, 4>(pl, st);
      default: return launch_pair<2, 2>(pl, st);
    }
  }
  if (pl->n_pad == 128) {
    if (pl->amode == kAModeSwizzle64) {
      if (pl->nacc == 4) return launch_variant<128, kAModeSwizzle64, 4, 1>(pl, st);
      switch (epi) {
        case 0: return launch_variant<128, kAModeSwizzle64, 2, 2, 0>(pl, st);
        case 1: return launch_variant<128, kAModeSwizzle64, 2, 2, 1>(pl, st);
        case 2: return launch_variant<128, kAModeSwizzle64, 2, 2, 2>(pl, st);
        case 3: return launch_variant<128, kAModeSwizzle64, 2, 2, 3>(pl, st);
        case 4: return launch_variant<128, kAModeSwizzle64, 2, 2, 4>(pl, st);
        default: return launch_variant<128, kAModeSwizzle64, 2, 2>(pl, st);
      }
    }
    return pl->nacc == 4 ? launch_variant<128, kAModeInterleave, 4, 1>(pl, st)
                         : launch_variant<128, kAModeInterleave, 2, 2>(pl, st);
  }
  if (pl->amode == kAModeSwizzle64) return launch_variant<16, kAModeSwizzle64, 4, 2>(pl, st);
  return launch_variant<16, kAModeInterleave, 4, 2>(pl, st);
}

extern "C" void sr_conv_plan_destroy(sr_conv_plan* plan) {
  delete reinterpret_cast<ConvPlan*>(plan);
}

extern "C" int sr_conv_plan_info(const sr_conv_plan* plan, sr_conv_plan_info_t* info) {
  if (!plan || !info) return set_error(SR_ERR_INVALID, "sr_conv_plan_info: null argument");
  const ConvPlan* pl = reinterpret_cast<const ConvPlan*>(plan);
  info->flops = pl->flops;
  info->total_tiles = pl->P.total_tiles;
  info->grid = pl->grid;
  info->smem_bytes = (int)pl->smem_bytes;
  info->seg_width = pl->P.BW;
  info->nseg = pl->P.nseg;
  info->strip_rows = pl->P.NR;
  info->num_wstages = pl->P.num_wstages;
  info->tile_positions = pl->nacc * 128;
  info->mma_efficiency =
      (double)pl->P.NB * pl->P.Hc * pl->P.Wc / ((double)pl->P.total_tiles * pl->nacc * 128);
  return SR_OK;
}


// ------------------------------------------------------------------ chain plan (see conv_tc_chain_kernel)
struct ConvChain {
  sr::ChainConv* dev = nullptr;       // device copy of the conv list
  unsigned* counter = nullptr;        // grid barrier arrivals (zeroed by every run)
  int n_convs = 0, n_phases = 0, grid = 0, a_slot_bytes = 0, num_abuf = 0, num_wstages = 0;
  size_t smem_bytes = 0;
  double flops = 0;
  int total_tiles = 0;
};

extern "C" int sr_conv_chain_create(const sr_conv_desc* descs, const int* phase, int n, sr_conv_chain** out) {
  if (!descs || !phase || !out || n < 1) return set_error(SR_ERR_INVALID, "sr_conv_chain_create: bad argument");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  const int wstage = kTapsPerStage * 64 * kRowBytes;      // a CTA's half of a two-tap weight stage
  const int stage_bytes = 4 * 32 * 256;
  std::vector<sr::ChainConv> host((size_t)n);
  for (int i = 0; i < n; ++i) {
    const sr_conv_desc* d = &descs[i];
    if (d->nsrc < 1 || d->nsrc > 2 || d->cin != kCin || d->cout != 128 || d->precision != 0 || d->a_mode != 0 ||
        d->shuffle_r > 0 || d->relu_mask_bf16 || d->colsum_f32 || d->relu == 2 || d->out_index || d->stitch_u8 ||
        d->cin_valid[0] || d->cin_valid[1] || !d->out_bf16 || (d->res_bf16 && !d->res_f32) || (d->out_f32 && !d->res_f32) || d->NB < 1 || d->H < 1 || d->W < 1)
      return set_error(SR_ERR_UNSUPPORTED, "sr_conv_chain_create: 128 -> 128 bf16 convs with the plain (bf16 out) or "
                                           "fp32-residual epilogue only");
    if (i > 0 && (phase[i] < phase[i - 1] || phase[i] > phase[i - 1] + 1))
      return set_error(SR_ERR_INVALID, "sr_conv_chain_create: phases must be non-decreasing without gaps");
    if (i == 0 && phase[0] != 0) return set_error(SR_ERR_INVALID, "sr_conv_chain_create: the first phase is 0");
    for (int s = 0; s < d->nsrc; ++s) {
      const int k = d->ksize[s];
      if (!(k == 1 || k == 3 || k == 5 || k == 7) || !d->in[s] || !d->wpacked[s])
        return set_error(SR_ERR_INVALID, "sr_conv_chain_create: bad source");
    }
  }
  // one strip-ring / weight-ring geometry for all convs: as many strip buffers as leave a useful weight ring
  int nab = 4, a_slot = 0, nstages = 0;
  for (; nab >= 2; --nab) {
    a_slot = 0;
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
      const sr_conv_desc* d = &descs[i];
      ConvKernelParams& P = host[i].P;
      memset(&P, 0, sizeof P);
      int p = 0;
      for (int s = 0; s < d->nsrc; ++s) p = std::max(p, (d->ksize[s] - 1) / 2);
      P.nsrc = d->nsrc;
      P.ksize[0] = d->ksize[0];
      P.ksize[1] = d->nsrc > 1 ? d->ksize[1] : 0;
      P.H = d->H, P.W = d->W, P.NB = d->NB, P.p = p;
      P.Hc = (d->comp_h > 0 && d->comp_h < d->H) ? d->comp_h : d->H;
      P.Wc = (d->comp_w > 0 && d->comp_w < d->W) ? d->comp_w : d->W;
      const int step = d->NB >= 2 ? 1 : 2;     // a single image: an even number of column segments (no idle CTA)
      ok = choose_geometry(P.Hc, P.Wc, p, 128, wstage, stage_bytes, &P, step, step, nab);
      a_slot = std::max(a_slot, P.a_bytes);
    }
    if (!ok) continue;
    const size_t fixed = 1024 + (size_t)nab * a_slot + stage_bytes + sizeof(ConvBarriers) + 4 * 128 * 4 + 64;
    if (fixed + 6 * (size_t)wstage > kSmemBudget && nab > 2) continue;
    if (fixed + 3 * (size_t)wstage > kSmemBudget) continue;
    nstages = std::min((int)((kSmemBudget - fixed) / wstage), kMaxWStages);
    break;
  }
  if (nab < 2 || nstages < 3) return set_error(SR_ERR_UNSUPPORTED, "sr_conv_chain_create: no geometry fits shared memory");
  ConvChain* ch = new (std::nothrow) ConvChain();
  if (!ch) return set_error(SR_ERR_NOMEM, "out of host memory");
  int max_phase_tiles = 0, cur_phase_tiles = 0;
  for (int i = 0; i < n; ++i) {
    const sr_conv_desc* d = &descs[i];
    sr::ChainConv& cv = host[i];
    ConvKernelParams& P = cv.P;
    P.num_abuf = nab;
    P.num_wstages = nstages;
    P.total_tiles = P.NB * P.nseg * P.tiles_per_seg;
    P.bias = d->bias;
    P.alpha = d->alpha, P.beta = d->beta, P.relu = d->relu;
    P.res_f32 = d->res_f32;
    P.out_bf16 = reinterpret_cast<__nv_bfloat16*>(d->out_bf16);
    P.out_f32 = d->out_f32;
    P.cout = d->cout;
    cv.phase = phase[i];
    cv.pair_tiles = ((P.NB * P.nseg + 1) / 2) * P.tiles_per_seg;
    if (i > 0 && phase[i] != phase[i - 1]) cur_phase_tiles = 0;
    cur_phase_tiles += cv.pair_tiles;
    max_phase_tiles = std::max(max_phase_tiles, cur_phase_tiles);
    ch->total_tiles += P.total_tiles;
    for (int s = 0; s < d->nsrc; ++s) {
      int rc = make_a_map(&cv.tmA[s], d->in[s], d->NB, d->H, d->W, P.PWs, P.NR, kAModeSwizzle64, 0);
      if (rc == SR_OK) rc = make_w_map(&cv.tmW[s], d->wpacked[s], kNumChunks * d->ksize[s] * d->ksize[s], 128, 64, 0);
      if (rc != SR_OK) {
        delete ch;
        return rc;
      }
      ch->flops += 2.0 * (double)d->NB * P.Hc * P.Wc * d->ksize[s] * d->ksize[s] * kCin * d->cout;
    }
    if (d->nsrc == 1) cv.tmA[1] = cv.tmA[0], cv.tmW[1] = cv.tmW[0];
  }
  ch->n_convs = n;
  ch->n_phases = phase[n - 1] + 1;
  ch->a_slot_bytes = a_slot, ch->num_abuf = nab, ch->num_wstages = nstages;
  ch->grid = 2 * std::min(max_phase_tiles, sms / 2);
  ch->smem_bytes = 1024 + (size_t)nab * a_slot + (size_t)nstages * wstage + stage_bytes + sizeof(ConvBarriers) +
                   4 * 128 * 4 + 64;
  void* devp = nullptr;
  if (cudaMalloc(&devp, (size_t)n * sizeof(sr::ChainConv) + 256) != cudaSuccess) {
    cudaGetLastError();
    delete ch;
    return set_error(SR_ERR_NOMEM, "sr_conv_chain_create: cudaMalloc failed");
  }
  ch->counter = reinterpret_cast<unsigned*>(devp);
  ch->dev = reinterpret_cast<sr::ChainConv*>(reinterpret_cast<char*>(devp) + 256);
  if (cudaMemcpy(ch->dev, host.data(), (size_t)n * sizeof(sr::ChainConv), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(devp);
    delete ch;
    return set_error(SR_ERR_CUDA, "sr_conv_chain_create: upload failed");
  }
  *out = reinterpret_cast<sr_conv_chain*>(ch);
  return SR_OK;
}

extern "C" int sr_conv_chain_run(sr_conv_chain* chain, void* stream) {
  if (!chain) return set_error(SR_ERR_INVALID, "sr_conv_chain_run: null chain");
  const ConvChain* ch = reinterpret_cast<const ConvChain*>(chain);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static unsigned long long attr_done = 0;
  if (int rc = ensure_dynamic_smem(conv_tc_chain_kernel, (int)kSmemBudget, &attr_done,
                                   "cudaFuncSetAttribute(conv_tc_chain_kernel)"))
    return rc;
  cudaError_t e = cudaMemsetAsync(ch->counter, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return set_cuda_error(e, "sr_conv_chain_run: memset");
  conv_tc_chain_kernel<<<ch->grid, kConvThreads, ch->smem_bytes, st>>>(ch->dev, ch->n_convs, ch->n_phases, ch->counter,
                                                                     ch->a_slot_bytes, ch->num_abuf, ch->num_wstages,
                                                                     dev_timeline());
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "conv_tc_chain_kernel launch");
  return SR_OK;
}

extern "C" void sr_conv_chain_destroy(sr_conv_chain* chain) {
  ConvChain* ch = reinterpret_cast<ConvChain*>(chain);
  if (!ch) return;
  if (ch->counter) cudaFree(ch->counter);
  delete ch;
}

extern "C" int sr_conv_chain_info(const sr_conv_chain* chain, sr_conv_plan_info_t* info) {
  if (!chain || !info) return set_error(SR_ERR_INVALID, "sr_conv_chain_info: null argument");
  const ConvChain* ch = reinterpret_cast<const ConvChain*>(chain);
  memset(info, 0, sizeof *info);
  info->flops = ch->flops;
  info->total_tiles = ch->total_tiles;
  info->grid = ch->grid;
  info->smem_bytes = (int)ch->smem_bytes;
  info->num_wstages = ch->num_wstages;
  info->tile_positions = 128;
  info->strip_rows = ch->n_phases;     // (reused field: number of phases)
  info->nseg = ch->n_convs;            // (reused field: number of convolutions)
  return SR_OK;
}
