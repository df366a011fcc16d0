// Y-channel PSNR + SSIM scoring (reference: scorpath.py:174-228, PSNR.py:54-84, skimage
// rgb2ycbcr / compare_ssim semantics restated in oracle/scoring.py).
// One fused pass: uint8 RGB pair -> crop -> integer {R,G,B,Z} planes in shared memory -> separable sliding 7x7
// window sums (exact integers; a block walks down a band of columns carrying the vertical sums in registers) ->
// SSIM map value -> warp reduction (shuffles) -> 64-bit integer atomics (bit-reproducible); see score_pair_kernel.
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.h"

namespace sr {
namespace {

constexpr int kWin = 7;
constexpr int kBW = 32;                     // window columns (= owned pixel columns) of a block's column band
constexpr int kPC = kBW + kWin - 1;         // 38 pixel columns staged per row
constexpr int kRS = 8;                      // pixel rows per step of the walk down the band
constexpr int kPxStride = 68;               // words per staged pixel row (permuted columns; == 4 mod 32)
constexpr int kHsStride = 36;               // words per row of horizontal sums (permuted columns; == 4 mod 32)
constexpr int kSlots = 16;                  // ring of row-sum rows: 8 new + 7 old
constexpr int kNVal = 15;                   // words per (row, column): 3 x {s, q, d2} + Y {sa, sb, q lo/hi, d2 lo/hi}
constexpr int kScoreThreads = 128;          // four warps = four roles (R, G, B, Y), rotated with the block index
constexpr unsigned kK = 255000u;            // Y = 16 + Z / kK with Z = 65481 R + 128553 G + 24966 B (integer)
constexpr double kFix = 1099511627776.0;    // SSIM values are accumulated as 2^-40 fixed point (int64)
constexpr unsigned long long kFixBias = 0x4338000000000000ull;   // bits of 1.5 * 2^52

__device__ __forceinline__ double y_from_rgb(double r, double g, double b) {
  // skimage.color.rgb2ycbcr on img_as_float(uint8): (r*65.481 + g*128.553 + b*24.966) + 16
  return ((r / 255.0) * 65.481 + (g / 255.0) * 128.553 + (b / 255.0) * 24.966) + 16.0;
}

__device__ __forceinline__ double block_sum(double v, double* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (wid == 0) {
    t = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

// bits of (1.5 * 2^52 + rint(2^40 * num / den)) for den > 0, |num / den| <= 1.
//   * the reciprocal is the hardware seed r0 (MUFU.RCP64H, relative error e with |e| < 2^-20) corrected
//     SR_SCORE_DIV_TERMS times: num / den = num r0 / (1 - e) = num r0 (1 + e + e^2 + ...), e = 1 - den r0; two
//     terms leave a relative error e^3 < 2^-60, below the fp64 rounding of the operands.  The chain is
//     seed -> e -> (e + e^2) -> quotient; num r0 does not wait for e.
//   * the scaling by 2^40 is an exponent increment of num (no multiplication); the integer then sits in the
//     mantissa of the sum with 1.5 * 2^52, rounded to nearest even by that addition.
#ifndef SR_SCORE_BLOCKS_PER_SM
#define SR_SCORE_BLOCKS_PER_SM 4   // blocks the launch aims at per SM (= the resident blocks: one wave)
#endif
#ifndef SR_SCORE_DIV_TERMS
#define SR_SCORE_DIV_TERMS 2
#endif
__device__ __forceinline__ unsigned long long fix40_div_bits(double num, double den) {
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(den));
  const double ns = __hiloint2double(__double2hiint(num) + (40 << 20), __double2loint(num));   // num * 2^40 (num != 0)
  const double t = ns * r0;
  const double e = fma(-den, r0, 1.0);
  const double p = SR_SCORE_DIV_TERMS >= 2 ? fma(e, e, e) : e;
  return (unsigned long long)__double_as_longlong(fma(t, p, t) + 6755399441055744.0);
}

// column permutations that make every shared-memory access of the kernel bank-conflict free (see the kernel)
__device__ __forceinline__ int px_col(int c) { return ((c & 7) << 2) + ((c >> 3) & 3) + ((c >> 5) << 5); }
__device__ __forceinline__ int hs_col(int x) { return ((x & 7) << 2) + (x >> 3); }

struct ScoreSmem {
  uint32_t px[5][kRS][kPxStride];          // R, G, B as (a | b << 16); Za; Zb -- the 8 pixel rows of the current step
  uint32_t hs[kNVal][kSlots][kHsStride];   // horizontal 7-sums of the per-pixel moments, ring of pixel rows
};

// SSIM of one window of a colour plane from its exact integer moments:
//   s = Sx | Sy << 16, q = sum (x^2 + y^2), d2 = sum (x - y)^2   (so 2 Sxy = q - d2)
// SSIM = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2)), u = S / 49, v = (49 S2 - S S) / (48 * 49).
// With both factors scaled by 49^2 and 48 * 49 the variable parts are the exact integers 2 Sx Sy, Sx^2 + Sy^2,
// 2 T = 49 (q - d2) - 2 Sx Sy and U = 49 q - Sx^2 - Sy^2 -- the cancellation that makes SSIM need fp64 happens in
// integers -- and only the four sums with the scaled constants and the ratio are fp64.
__device__ __forceinline__ unsigned long long ssim_colour_fix(uint32_t s, uint32_t q, uint32_t d2) {
  constexpr double k1 = 49.0 * 49.0 * 6.5025, k2 = 48.0 * 49.0 * 58.5225;      // 15612.5025, 137644.92
  const uint32_t Sx = s & 0xffffu, Sy = s >> 16;
  const uint32_t P2 = 2u * Sx * Sy, XY = Sx * Sx + Sy * Sy;
  const int T2 = (int)(49u * (q - d2)) - (int)P2;                // 2 * 48 * 49 * vxy
  const uint32_t U = 49u * q - XY;                               // 48 * 49 * (vx + vy) >= 0
  const double A1 = (double)P2 + k1, B1 = (double)XY + k1;
  const double A2 = (double)T2 + k2, B2 = (double)U + k2;
  return fix40_div_bits(A1 * A2, B1 * B2);
}

// The same for Y = 16 + Z / 255000 from the integer moments of Z (second moments need 64 bits):
//   sa = sum Za, sb = sum Zb, q = sum (Za^2 + Zb^2), d2 = sum (Za - Zb)^2.  49 q < 2^64 (Z < 2^25.74).
__device__ __forceinline__ unsigned long long ssim_y_fix(uint32_t sa, uint32_t sb, unsigned long long q,
                                                         unsigned long long d2) {
  const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
  const double inv49k = 1.0 / (49.0 * (double)kK);
  const double invv = 1.0 / (48.0 * 49.0 * (double)kK * (double)kK);
  const unsigned long long pa = (unsigned long long)sa * sa, pb = (unsigned long long)sb * sb;
  const unsigned long long U = 49ull * q - pa - pb;                                        // Txx + Tyy >= 0
  const long long T2 = (long long)(49ull * (q - d2) - 2ull * ((unsigned long long)sa * sb));  // 2 Txy
  const double ux = fma((double)sa, inv49k, 16.0), uy = fma((double)sb, inv49k, 16.0);
  const double A1 = fma(2.0 * ux, uy, C1), B1 = fma(ux, ux, fma(uy, uy, C1));
  const double A2 = fma((double)T2, invv, C2), B2 = fma((double)U, invv, C2);
  return fix40_div_bits(A1 * A2, B1 * B2);
}

// Fused scoring of one image pair in EXACT INTEGER window arithmetic (reference: scorpath.py:174-228).
//   * R, G, B are bytes and Y = 16 + Z / 255000 with Z an integer < 2^26, so every 7x7 window moment is an
//     integer; the variance / covariance numerators (49 S2 - S S) are formed in integers -- the cancellation that
//     makes SSIM need fp64 never meets a rounded number -- and only the final ratio is fp64.
//   * a block owns a band of 32 window columns and WALKS DOWN a chunk of rows, 8 pixel rows per step:
//       convert   128 threads turn the step's 8 x 38 pixels into packed planes in shared memory (the next step's
//                 bytes are fetched into registers one step ahead) and add the owned pixels' squared Y error;
//       H         one warp per plane (R, G, B, Y; the role rotates with the block index so that the heavier Y warps
//                 spread over the SM's four schedulers): lane = (row, 8-column segment) computes the per-pixel
//                 moments of 14 pixels once and slides the horizontal 7-sums along its 8 outputs;
//       V         lane = window column keeps the vertical 7-row sums in REGISTERS across the whole walk (add the new
//                 row, subtract the row seven up, both from a 16-row ring), forms SSIM and accumulates.
//     No halo rows are recomputed between steps; the only redundancy is 6 columns per band and 6 rows per chunk.
//   * three moments per colour plane (s, q = sum x^2 + y^2, d2 = sum (x - y)^2) instead of four.
//   * columns are stored permuted (px_col / hs_col) with row strides == 4 mod 32 words: every load and store of
//     the three phases touches 32 distinct banks.
//   * per-window SSIM values are accumulated as 2^-40 fixed point and all cross-block sums are 64-bit INTEGER
//     atomics, so the result is bit-reproducible run to run; the last block to finish (ticket counter) converts
//     the accumulators into the public double fields.
// One launch scores up to kScoreBatch pairs (blockIdx.z); the pairs may differ in shape, a block outside its pair's
// bands x chunks leaves at once.
constexpr int kScoreBatch = 32;
struct ScoreBatchParams {
  sr_score_item item[kScoreBatch];
  int chunk_rows[kScoreBatch], steps[kScoreBatch], bands[kScoreBatch], chunks[kScoreBatch];
};

__global__ void __launch_bounds__(kScoreThreads) score_pair_kernel(const __grid_constant__ ScoreBatchParams B, int crop,
                                                                   sr_score_result* __restrict__ results) {
  __shared__ ScoreSmem sm;
  const int bands = B.bands[blockIdx.z], chunks = B.chunks[blockIdx.z];
  if ((int)blockIdx.x >= bands || (int)blockIdx.y >= chunks) return;
  const uint8_t* __restrict__ a = B.item[blockIdx.z].a;
  const uint8_t* __restrict__ b = B.item[blockIdx.z].b;
  const int h = B.item[blockIdx.z].h, w = B.item[blockIdx.z].w;
  const int chunk_rows = B.chunk_rows[blockIdx.z], steps = B.steps[blockIdx.z];
  sr_score_result* __restrict__ res = results + blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31;
  const int role = ((tid >> 5) + blockIdx.x + blockIdx.y) & 3;      // 0..2 colour plane, 3 = Y
  const int ch_ = h - 2 * crop, cw_ = w - 2 * crop;                 // cropped size
  const int nwy = ch_ - kWin + 1, nwx = cw_ - kWin + 1;             // valid windows per axis
  const int bx0 = blockIdx.x * kBW, cy0 = blockIdx.y * chunk_rows;  // first owned pixel = first window of the block

  // ---- convert (the three colour warps; the Y warp's H and V passes are the longer ones): pixel k of a thread
  // is index ctid + 96 k of the step's 8 x 38 pixels.  The three bytes of a pixel come from the one or two aligned
  // 32-bit words that hold them (funnel shift); a step advances every pointer by 8 rows = 24 w bytes, a multiple
  // of four, so the alignment of a pixel is the same in every step.
  constexpr int kCvtThreads = 96, kPixPerThread = (kRS * kPC + kCvtThreads - 1) / kCvtThreads;   // 4 (the 4th: 16 threads)
  const int ctid = role * 32 + lane;                   // only used when role < 3
  uint32_t p_wa[kPixPerThread], p_wb[kPixPerThread];   // byte offset of the pixel's first aligned word from base_a / base_b
  uint32_t p_inf[kPixPerThread];                       // shift of a | shift of b << 8 | row in the step (255: none) << 16 | owned column << 24
  int p_so[kPixPerThread];                             // shared-memory word offset
  uint32_t raw[kPixPerThread][4];                      // prefetched words: a lo, a hi, b lo, b hi
  const uint8_t* blk_a = a + ((size_t)(cy0 + crop) * w + (bx0 + crop)) * 3;     // the block's first pixel, step 0
  const uint8_t* blk_b = b + ((size_t)(cy0 + crop) * w + (bx0 + crop)) * 3;
  const uint32_t da = (uint32_t)(reinterpret_cast<uintptr_t>(blk_a) & 3u), db = (uint32_t)(reinterpret_cast<uintptr_t>(blk_b) & 3u);
  const uint8_t* base_a = blk_a - da;                  // aligned down; advanced by 24 w bytes per step
  const uint8_t* base_b = blk_b - db;
#pragma unroll
  for (int k = 0; k < kPixPerThread; ++k) {
    const int idx = ctid + k * kCvtThreads;
    const int r = idx / kPC, c = idx - r * kPC;
    const bool on = role < 3 && r < kRS && bx0 + c < cw_;        // columns right of the image stay zero
    const uint32_t oa = da + (uint32_t)(r * w + c) * 3u, ob = db + (uint32_t)(r * w + c) * 3u;
    p_wa[k] = oa & ~3u;
    p_wb[k] = ob & ~3u;
    p_inf[k] = ((oa & 3u) * 8u) | (((ob & 3u) * 8u) << 8) | ((uint32_t)(on ? r : 255) << 16) | ((on && c < kBW ? 1u : 0u) << 24);
    p_so[k] = r * kPxStride + px_col(c);
    raw[k][0] = raw[k][1] = raw[k][2] = raw[k][3] = 0u;
  }
  const size_t step_bytes = (size_t)kRS * w * 3;
  // fetch only LOADS, straight into the registers convert() reads one step later (no other use, no merge with a
  // zero: a pixel below the image keeps the stale words of the step before -- such pixels only reach windows that
  // are not scored, and the squared error below is guarded by own_rows)
  auto fetch = [&](int step) {
    const int rows_left = ch_ - cy0 - step * kRS;      // image rows from the step's first row on
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
      const int r = (int)((p_inf[k] >> 16) & 255u);    // 255 for pixels this thread does not convert
      const bool ok = r < rows_left;
      const uint32_t* qa = reinterpret_cast<const uint32_t*>(base_a + p_wa[k]);
      const uint32_t* qb = reinterpret_cast<const uint32_t*>(base_b + p_wb[k]);
      if (ok) raw[k][0] = qa[0];
      if (ok && (p_inf[k] & 0x10u) != 0u) raw[k][1] = qa[1];      // shift 16 or 24: the bytes spill into the next word
      if (ok) raw[k][2] = qb[0];
      if (ok && (p_inf[k] & 0x1000u) != 0u) raw[k][3] = qb[1];
    }
    base_a += step_bytes;
    base_b += step_bytes;
  };
  unsigned long long ssd = 0;
  const int own_rows = min(chunk_rows, ch_ - cy0);     // pixel rows of the chunk this block owns
  auto convert = [&](int step) {
#pragma unroll
    for (int k = 0; k < kPixPerThread; ++k) {
      if (k < kPixPerThread - 1 || ctid < kRS * kPC - (kPixPerThread - 1) * kCvtThreads) {
        const uint32_t va = __funnelshift_r(raw[k][0], raw[k][1], p_inf[k] & 31u) & 0xffffffu;   // r | g << 8 | b << 16
        const uint32_t vb = __funnelshift_r(raw[k][2], raw[k][3], (p_inf[k] >> 8) & 31u) & 0xffffffu;
        const uint32_t ar = va & 0xffu, ag = (va >> 8) & 0xffu, ab = va >> 16;
        const uint32_t br = vb & 0xffu, bg = (vb >> 8) & 0xffu, bb = vb >> 16;
        const uint32_t za = 65481u * ar + 128553u * ag + 24966u * ab;
        const uint32_t zb = 65481u * br + 128553u * bg + 24966u * bb;
        uint32_t* o = &sm.px[0][0][0] + p_so[k];
        constexpr int kPl = kRS * kPxStride;
        o[0] = __byte_perm(va, vb, 0x3430);            // a.r | b.r << 16 (byte 3 of va is zero)
        o[kPl] = __byte_perm(va, vb, 0x3531);
        o[2 * kPl] = __byte_perm(va, vb, 0x3632);
        o[3 * kPl] = za;
        o[4 * kPl] = zb;
        // squared Y error of the pixels this block owns
        if ((p_inf[k] >> 24) != 0u && step * kRS + (int)((p_inf[k] >> 16) & 255u) < own_rows) {
          const int d = (int)za - (int)zb;
          ssd += (unsigned long long)((long long)d * d);
        }
      }
    }
  };

  // ---- per-lane state of the two sliding phases
  const int hr = lane >> 2, hseg = lane & 3;                    // H: pixel row of the step, 8-column segment
  const int h_in0 = hr * kPxStride + hseg;                      // px_col(8 seg + i)     = 4 i + seg        (i < 8)
  const int h_in1 = hr * kPxStride + (hseg == 3 ? 32 : hseg + 1);   // px_col(8 seg + 8 + i) = 4 i + this   (i < 6)
  const int h_out = hr * kHsStride + hseg;                      // hs_col(8 seg + j)     = 4 j + seg
  const int v_col = hs_col(lane);                               // V: window column = lane
  // window rows of the chunk this lane scores: [0, wend) (none for a column right of the last window)
  const int wend = bx0 + lane < nwx ? min(chunk_rows, nwy - cy0) : 0;
  uint32_t w0 = 0, w1 = 0, w2 = 0;                              // vertical sums: colour {s, q, d2}; Y {sa, sb} + 64-bit pair
  unsigned long long wq = 0, wd = 0;
  unsigned long long acc = 0;                                   // sum of fix40_div_bits (bias removed at the end)
  constexpr int kPlane = kSlots * kHsStride;

  if (role < 3) {
    fetch(0);
    convert(0);
  }
  __syncthreads();

  for (int step = 0; step < steps; ++step) {
    if (role < 3 && step + 1 < steps) fetch(step + 1);
    const int half = (step & 1) * kRS * kHsStride;              // the step's 8 rows: ring slots 0-7 or 8-15

    // ---- H: horizontal 7-sums of the step's 8 pixel rows
    if (role < 3) {
      const uint32_t* in = &sm.px[role][0][0];
      uint32_t ms[14], mq[14], md[14];
#pragma unroll
      for (int i = 0; i < 14; ++i) {
        const uint32_t p = i < 8 ? in[h_in0 + 4 * i] : in[h_in1 + 4 * (i - 8)];
        const uint32_t x = p & 0xffffu, y = p >> 16;
        const int d = (int)x - (int)y;
        ms[i] = p;
        mq[i] = x * x + y * y;
        md[i] = (uint32_t)(d * d);
      }
      uint32_t* o0 = &sm.hs[3 * role][0][0] + half + h_out;
      uint32_t s = 0, q = 0, d2 = 0;
#pragma unroll
      for (int i = 0; i < 7; ++i) { s += ms[i]; q += mq[i]; d2 += md[i]; }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j > 0) { s += ms[j + 6] - ms[j - 1]; q += mq[j + 6] - mq[j - 1]; d2 += md[j + 6] - md[j - 1]; }
        o0[4 * j] = s;
        o0[kPlane + 4 * j] = q;
        o0[2 * kPlane + 4 * j] = d2;
      }
    } else {
      const uint32_t* ina = &sm.px[3][0][0];
      const uint32_t* inb = &sm.px[4][0][0];
      uint32_t za[14], zb[14];
#pragma unroll
      for (int i = 0; i < 14; ++i) {
        const int o = i < 8 ? h_in0 + 4 * i : h_in1 + 4 * (i - 8);
        za[i] = ina[o];
        zb[i] = inb[o];
      }
      uint32_t* o0 = &sm.hs[9][0][0] + half + h_out;
      uint32_t sa = 0, sb = 0;
      unsigned long long q = 0, d2 = 0;
      auto mom_q = [](uint32_t x, uint32_t y) { return (unsigned long long)x * x + (unsigned long long)y * y; };
      auto mom_d = [](uint32_t x, uint32_t y) {
        const int d = (int)x - (int)y;
        return (unsigned long long)((long long)d * d);
      };
#pragma unroll
      for (int i = 0; i < 7; ++i) { sa += za[i]; sb += zb[i]; q += mom_q(za[i], zb[i]); d2 += mom_d(za[i], zb[i]); }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j > 0) {
          sa += za[j + 6] - za[j - 1];
          sb += zb[j + 6] - zb[j - 1];
          q += mom_q(za[j + 6], zb[j + 6]) - mom_q(za[j - 1], zb[j - 1]);
          d2 += mom_d(za[j + 6], zb[j + 6]) - mom_d(za[j - 1], zb[j - 1]);
        }
        o0[4 * j] = sa;
        o0[kPlane + 4 * j] = sb;
        o0[2 * kPlane + 4 * j] = (uint32_t)q;
        o0[3 * kPlane + 4 * j] = (uint32_t)(q >> 32);
        o0[4 * kPlane + 4 * j] = (uint32_t)d2;
        o0[5 * kPlane + 4 * j] = (uint32_t)(d2 >> 32);
      }
    }
    __syncthreads();                         // row sums of the step are in the ring; the pixel buffer is free
    if (role < 3 && step + 1 < steps) convert(step + 1);

    // ---- V: vertical 7-row sums carried in registers, SSIM of the window rows that complete in this step.
    // Pixel row j of the step is ring row half + j; the row seven up is row j + 1 of the OTHER half (j < 7) or
    // row 0 of this half (j = 7): every address is a base plus an immediate.
    {
      const int wrow0 = step * kRS - (kWin - 1);   // window row completed by the step's first pixel row
      const bool first = step == 0;                // rows 0..6 of the chunk have nothing to subtract / complete
      if (role < 3) {
        const uint32_t* cur = &sm.hs[3 * role][0][0] + half + v_col;
        const uint32_t* oth = &sm.hs[3 * role][0][0] + (half ^ (kRS * kHsStride)) + v_col;
#pragma unroll
        for (int j = 0; j < kRS; ++j) {
          const uint32_t* n = cur + j * kHsStride;
          w0 += n[0]; w1 += n[kPlane]; w2 += n[2 * kPlane];
          if (j == kRS - 1 || !first) {
            const uint32_t* o = j == kRS - 1 ? cur : oth + (j + 1) * kHsStride;
            w0 -= o[0]; w1 -= o[kPlane]; w2 -= o[2 * kPlane];
          }
          if ((j >= kWin - 1 || !first) && wrow0 + j < wend) acc += ssim_colour_fix(w0, w1, w2);
        }
      } else {
        const uint32_t* cur = &sm.hs[9][0][0] + half + v_col;
        const uint32_t* oth = &sm.hs[9][0][0] + (half ^ (kRS * kHsStride)) + v_col;
#pragma unroll
        for (int j = 0; j < kRS; ++j) {
          const uint32_t* n = cur + j * kHsStride;
          w0 += n[0]; w1 += n[kPlane];
          wq += (unsigned long long)n[2 * kPlane] | ((unsigned long long)n[3 * kPlane] << 32);
          wd += (unsigned long long)n[4 * kPlane] | ((unsigned long long)n[5 * kPlane] << 32);
          if (j == kRS - 1 || !first) {
            const uint32_t* o = j == kRS - 1 ? cur : oth + (j + 1) * kHsStride;
            w0 -= o[0]; w1 -= o[kPlane];
            wq -= (unsigned long long)o[2 * kPlane] | ((unsigned long long)o[3 * kPlane] << 32);
            wd -= (unsigned long long)o[4 * kPlane] | ((unsigned long long)o[5 * kPlane] << 32);
          }
          if ((j >= kWin - 1 || !first) && wrow0 + j < wend) acc += ssim_y_fix(w0, w1, wq, wd);
        }
      }
    }
    __syncthreads();                         // the next step's pixels are staged; ring rows older than 7 are free
  }

  // ---- reduction: integers, so the order does not matter.  One atomic per warp and accumulator, then the last
  // block (ticket) converts.
  acc -= (unsigned long long)(wend > 0 ? wend : 0) * kFixBias;   // every scored window added the bias once
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    ssd += __shfl_xor_sync(0xffffffffu, ssd, o);
  }
  unsigned long long* gacc = reinterpret_cast<unsigned long long*>(res->acc);
  if (lane == 0) {
    atomicAdd(gacc + (role == 3 ? 0 : 1 + role), acc);
    atomicAdd(gacc + 4, ssd & 0xffffffffull);      // squared error: low / high halves (the total can exceed 2^64)
    atomicAdd(gacc + 5, ssd >> 32);
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(&res->ticket), 1ull);
    if (done == (unsigned long long)bands * chunks - 1) {
      __threadfence();
      volatile unsigned long long* ac = gacc;
      const double kk = (double)kK * (double)kK;
      res->ssim_y_sum = (double)(long long)ac[0] / kFix;
      for (int k = 0; k < 3; ++k) res->ssim_rgb_sum[k] = (double)(long long)ac[1 + k] / kFix;
      res->sum_sq_y = ((double)ac[4] + (double)ac[5] * 4294967296.0) / kk;
      res->n_pix = (int64_t)ch_ * cw_;
      res->n_win = (int64_t)(nwy > 0 ? nwy : 0) * (nwx > 0 ? nwx : 0);
    }
  }
}

__global__ void rgb2y_kernel(const uint8_t* __restrict__ rgb, size_t npix, double* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] = y_from_rgb(rgb[i * 3], rgb[i * 3 + 1], rgb[i * 3 + 2]);
}

__global__ void sum_sq_diff_f64_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                       size_t n, double* __restrict__ out) {
  __shared__ double red[32];
  double local = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const double d = a[i] - b[i];
    local += d * d;
  }
  const double t = block_sum(local, red);
  if (threadIdx.x == 0) atomicAdd(out, t);
}

}  // namespace
}  // namespace sr

using namespace sr;

extern "C" int sr_sum_sq_diff_f64(const double* a, const double* b, size_t n, double* out, void* stream) {
  if (!a || !b || !out) return set_error(SR_ERR_INVALID, "sr_sum_sq_diff_f64: null pointer");
  if (n == 0) return SR_OK;
  sum_sq_diff_f64_kernel<<<grid_for(n, 256, 148 * 4), 256, 0, as_stream(stream)>>>(a, b, n, out);
  return check_launch("sum_sq_diff_f64_kernel");
}

extern "C" int sr_rgb2y_u8(const uint8_t* rgb, size_t npix, double* y, void* stream) {
  if (!rgb || !y) return set_error(SR_ERR_INVALID, "sr_rgb2y_u8: null pointer");
  if (npix == 0) return SR_OK;
  rgb2y_kernel<<<grid_for(npix, 256), 256, 0, as_stream(stream)>>>(rgb, npix, y);
  return check_launch("rgb2y_kernel");
}

// geometry of one pair inside a launch that aims at `blocks_wanted` blocks for it: bands of 32 columns x chunks of
// rows (chunks of at least 26 window rows on a small image); a chunk is walked in steps of 8 pixel rows (chunk + 6
// halo rows)
static void score_geometry(int h, int w, int crop, int blocks_wanted, int* bands, int* chunks, int* chunk_rows, int* steps) {
  const int ch_ = h - 2 * crop, cw_ = w - 2 * crop;
  *bands = (cw_ + kBW - 1) / kBW;
  int nch = (blocks_wanted + *bands / 2) / *bands;
  const int max_chunks = (ch_ + 25) / 26;
  if (nch > max_chunks) nch = max_chunks;
  if (nch < 1) nch = 1;
  int rows = (ch_ + nch - 1) / nch;
  *steps = (rows + kWin - 1 + kRS - 1) / kRS;
  *chunk_rows = *steps * kRS - (kWin - 1);
  *chunks = (ch_ + *chunk_rows - 1) / *chunk_rows;
}

extern "C" int sr_score_batch_u8(const sr_score_item* items, int n, int crop, sr_score_result* results, void* stream) {
  if (n < 0 || (n > 0 && (!items || !results))) return set_error(SR_ERR_INVALID, "sr_score_batch_u8: null pointer");
  for (int i = 0; i < n; ++i) {
    if (!items[i].a || !items[i].b) return set_error(SR_ERR_INVALID, "sr_score_batch_u8: null image pointer");
    if (crop < 0 || items[i].h - 2 * crop < kWin || items[i].w - 2 * crop < kWin)
      return set_error(SR_ERR_INVALID, "sr_score_pair_u8: image smaller than the 7x7 SSIM window after cropping");
  }
  for (int i0 = 0; i0 < n; i0 += kScoreBatch) {
    const int m = n - i0 < kScoreBatch ? n - i0 : kScoreBatch;
    // about one resident wave of blocks (SR_SCORE_BLOCKS_PER_SM per SM) over the whole launch
    const int per_item = (148 * SR_SCORE_BLOCKS_PER_SM + m - 1) / m;
    ScoreBatchParams B = {};
    unsigned gx = 1, gy = 1;
    for (int i = 0; i < m; ++i) {
      B.item[i] = items[i0 + i];
      score_geometry(items[i0 + i].h, items[i0 + i].w, crop, per_item, &B.bands[i], &B.chunks[i], &B.chunk_rows[i], &B.steps[i]);
      if ((unsigned)B.bands[i] > gx) gx = (unsigned)B.bands[i];
      if ((unsigned)B.chunks[i] > gy) gy = (unsigned)B.chunks[i];
    }
    if (gy > 65535u) return set_error(SR_ERR_UNSUPPORTED, "sr_score_pair_u8: image taller than 1.7 M rows");
    score_pair_kernel<<<dim3(gx, gy, (unsigned)m), kScoreThreads, 0, as_stream(stream)>>>(B, crop, results + i0);
    if (int rc = check_launch("score_pair_kernel")) return rc;
  }
  return SR_OK;
}

extern "C" int sr_score_pair_u8(const uint8_t* a, const uint8_t* b, int h, int w, int crop,
                                sr_score_result* result, void* stream) {
  if (!a || !b || !result) return set_error(SR_ERR_INVALID, "sr_score_pair_u8: null pointer");
  const sr_score_item item = {a, b, h, w};
  return sr_score_batch_u8(&item, 1, crop, result, stream);
}
