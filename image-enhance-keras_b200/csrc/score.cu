// Y-channel PSNR + SSIM scoring (reference: scorpath.py:174-228, PSNR.py:54-84, skimage
// rgb2ycbcr / compare_ssim semantics restated in oracle/scoring.py).
// One fused pass: uint8 RGB pair -> crop -> integer {R,G,B,Z} planes in shared memory -> separable sliding 7x7
// window sums (exact integers) -> SSIM map value -> block reduction (warp shuffles) -> 64-bit integer atomics
// (bit-reproducible); see score_pair_kernel.
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.h"

namespace sr {
namespace {

constexpr int kWin = 7;
constexpr int kTW = 32, kTH = 26;                       // window positions (= cropped pixels owned) per block
constexpr int kHW = kTW + kWin - 1, kHH = kTH + kWin - 1;  // 38 x 32 halo pixels: 32 rows x 8 column runs = 256 threads
constexpr int kNPix = kHH * kHW, kNH = kHH * kTW;
constexpr unsigned kK = 255000u;                        // Y = 16 + Z / kK with Z = 65481 R + 128553 G + 24966 B (integer)
constexpr int kScoreSmem = (5 * kNPix + 8 * kNH) * 4;   // 24,320 B of pixels + 32,768 B of row sums
constexpr double kFix = 1099511627776.0;                // SSIM values are accumulated as 2^-40 fixed point (int64)

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

// a / b for b > 0: fp32 reciprocal refined by two Newton steps in fp64 (relative error ~1e-16; a full-precision
// fp64 division costs several times as much and this kernel does four per pixel)
__device__ __forceinline__ double fast_div(double a, double b) {
  double r = (double)__frcp_rn((float)b);
  r = r * (2.0 - b * r);
  r = r * (2.0 - b * r);
  return a * r;
}

// Fused scoring of one image pair in EXACT INTEGER window arithmetic.
//   * R, G, B are bytes and Y = 16 + Z / 255000 with Z an integer < 2^26, so every 7x7 window moment (sum x, sum y,
//     sum x^2, sum y^2, sum xy) is an integer: < 2^32 for the colour planes, < 2^63 for Z -- no rounding, any order.
//   * separable SLIDING sums: one pass of horizontal 7-sums per halo row (a thread slides along 4 columns), one
//     pass of vertical 7-sums (a thread slides down 4 windows); ~3 pixel evaluations per window and pass instead of
//     the 17 of summing taps directly.
//   * the variance / covariance numerators (49 Sxy - Sx Sy, ...) are formed in integers -- the cancellation that
//     makes SSIM need fp64 never meets a rounded number -- then four conversions, two products and one division
//     per plane in fp64.
//   * per-window SSIM values are accumulated as 2^-40 fixed point and all cross-block sums are 64-bit INTEGER
//     atomics, so the result is bit-reproducible run to run (fp64 atomics were not); the last block to finish
//     (ticket counter) converts the accumulators into the public double fields.
__global__ void __launch_bounds__(256) score_pair_kernel(const uint8_t* __restrict__ a,
                                                         const uint8_t* __restrict__ b, int h, int w,
                                                         int crop, sr_score_result* __restrict__ res) {
  extern __shared__ uint32_t sm[];
  uint32_t* px = sm;                 // [5][kNPix]: R, G, B as (a | b << 16); Za; Zb
  uint32_t* hb = sm + 5 * kNPix;     // row sums: colour phase 4 x u32 [kNH]; Y phase 2 x u32 + 3 x u64 [kNH]
  __shared__ long long red[8][5];
  const int tid = threadIdx.x;
  const int ch_ = h - 2 * crop, cw_ = w - 2 * crop;                 // cropped size
  const int wy0 = blockIdx.y * kTH, wx0 = blockIdx.x * kTW;         // first window (top-left) / first owned pixel
  const int nwy = ch_ - kWin + 1, nwx = cw_ - kWin + 1;             // valid windows per axis

  // ---- pixels -> shared memory; squared Y error of the owned pixels (exact: (Za - Zb)^2)
  unsigned long long ssd = 0;
  for (int i = tid; i < kNPix; i += 256) {
    const int ly = i / kHW, lx = i - ly * kHW;
    const int y = wy0 + ly, x = wx0 + lx;
    uint32_t pr = 0, pg = 0, pb = 0, za = 0, zb = 0;
    if (y < ch_ && x < cw_) {
      const size_t o = ((size_t)(y + crop) * w + (x + crop)) * 3;
      const uint32_t ar = a[o], ag = a[o + 1], ab = a[o + 2], br = b[o], bg = b[o + 1], bb = b[o + 2];
      pr = ar | (br << 16);
      pg = ag | (bg << 16);
      pb = ab | (bb << 16);
      za = 65481u * ar + 128553u * ag + 24966u * ab;
      zb = 65481u * br + 128553u * bg + 24966u * bb;
      if (ly < kTH && lx < kTW) {
        const long long d = (long long)za - (long long)zb;
        ssd += (unsigned long long)(d * d);
      }
    }
    px[i] = pr;
    px[kNPix + i] = pg;
    px[2 * kNPix + i] = pb;
    px[3 * kNPix + i] = za;
    px[4 * kNPix + i] = zb;
  }
  __syncthreads();

  const int hr = tid >> 3, hx0 = (tid & 7) * 4;           // horizontal task: halo row, first of 4 window columns
  const int vx = tid & 31, vy0 = (tid >> 5) * 4;          // vertical task: window column, first of <= 4 window rows
  const int vrows = min(4, kTH - vy0);
  const bool col_ok = wx0 + vx < nwx;
  long long acc_y = 0, acc_r = 0, acc_g = 0, acc_b = 0;   // fixed-point SSIM sums

  // ---- colour planes
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    long long cacc = 0;
    {
      const uint32_t* row = px + c * kNPix + hr * kHW + hx0;
      uint32_t s1 = 0, sxx = 0, syy = 0, sxy = 0;         // s1 = sum x | sum y << 16 (each <= 7 * 255)
#pragma unroll
      for (int dx = 0; dx < kWin; ++dx) {
        const uint32_t p = row[dx], x = p & 0xffffu, y = p >> 16;
        s1 += p; sxx += x * x; syy += y * y; sxy += x * y;
      }
      uint32_t* o = hb + hr * kTW + hx0;
      o[0] = s1; o[kNH] = sxx; o[2 * kNH] = syy; o[3 * kNH] = sxy;
#pragma unroll
      for (int j = 1; j < 4; ++j) {
        const uint32_t pn = row[j + kWin - 1], xn = pn & 0xffffu, yn = pn >> 16;
        const uint32_t po = row[j - 1], xo = po & 0xffffu, yo = po >> 16;
        s1 += pn - po; sxx += xn * xn - xo * xo; syy += yn * yn - yo * yo; sxy += xn * yn - xo * yo;
        o[j] = s1; o[kNH + j] = sxx; o[2 * kNH + j] = syy; o[3 * kNH + j] = sxy;
      }
    }
    __syncthreads();
    if (vrows > 0) {
      const uint32_t* col = hb + vy0 * kTW + vx;
      uint32_t s1 = 0, sxx = 0, syy = 0, sxy = 0;         // window sums: s1 fields <= 49 * 255 = 12495
#pragma unroll
      for (int dy = 0; dy < kWin; ++dy) {
        const uint32_t* q = col + dy * kTW;
        s1 += q[0]; sxx += q[kNH]; syy += q[2 * kNH]; sxy += q[3 * kNH];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < vrows) {
          if (j > 0) {
            const uint32_t* qn = col + (j + kWin - 1) * kTW;
            const uint32_t* qo = col + (j - 1) * kTW;
            s1 += qn[0] - qo[0]; sxx += qn[kNH] - qo[kNH]; syy += qn[2 * kNH] - qo[2 * kNH];
            sxy += qn[3 * kNH] - qo[3 * kNH];
          }
          if (col_ok && wy0 + vy0 + j < nwy) {
            // SSIM = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2)) with u = S / 49, v = 49/48 (..):
            // numerator and denominator scaled by 10^4 * 49^2 (first factor) and 100 * 48 * 49 (second) are integers
            const uint32_t Sx = s1 & 0xffffu, Sy = s1 >> 16;
            const uint32_t P = Sx * Sy, X2 = Sx * Sx, Y2 = Sy * Sy;
            const long long NA1 = 20000ll * P + 156125025ll;            // 10^4 (2 Sx Sy + 49^2 C1), C1 = 6.5025
            const long long NB1 = 10000ll * ((long long)X2 + Y2) + 156125025ll;
            const int T = (int)(49u * sxy) - (int)P;                      // 48 * 49 * vxy
            const uint32_t U = 49u * (sxx + syy) - X2 - Y2;               // 48 * 49 * (vx + vy) >= 0
            const long long NA2 = 200ll * T + 13764492ll;                 // 100 (2 T + 2352 C2), C2 = 58.5225
            const long long NB2 = 100ll * U + 13764492ll;
            const double v = fast_div((double)NA1 * (double)NA2, (double)NB1 * (double)NB2);
            cacc += __double2ll_rn(v * kFix);
          }
        }
      }
    }
    acc_r += c == 0 ? cacc : 0;
    acc_g += c == 1 ? cacc : 0;
    acc_b += c == 2 ? cacc : 0;
    __syncthreads();
  }

  // ---- Y plane: the same two passes on Z (sums of Z fit u32, second moments need u64)
  {
    unsigned long long* hq = reinterpret_cast<unsigned long long*>(hb + 2 * kNH);   // [3][kNH]
    {
      const uint32_t* ra = px + 3 * kNPix + hr * kHW + hx0;
      const uint32_t* rb = px + 4 * kNPix + hr * kHW + hx0;
      uint32_t sa = 0, sb = 0;
      unsigned long long saa = 0, sbb = 0, sab = 0;
#pragma unroll
      for (int dx = 0; dx < kWin; ++dx) {
        const unsigned long long x = ra[dx], y = rb[dx];
        sa += (uint32_t)x; sb += (uint32_t)y; saa += x * x; sbb += y * y; sab += x * y;
      }
      const int o = hr * kTW + hx0;
      hb[o] = sa; hb[kNH + o] = sb; hq[o] = saa; hq[kNH + o] = sbb; hq[2 * kNH + o] = sab;
#pragma unroll
      for (int j = 1; j < 4; ++j) {
        const unsigned long long xn = ra[j + kWin - 1], yn = rb[j + kWin - 1], xo = ra[j - 1], yo = rb[j - 1];
        sa += (uint32_t)xn - (uint32_t)xo; sb += (uint32_t)yn - (uint32_t)yo;
        saa += xn * xn - xo * xo; sbb += yn * yn - yo * yo; sab += xn * yn - xo * yo;
        hb[o + j] = sa; hb[kNH + o + j] = sb; hq[o + j] = saa; hq[kNH + o + j] = sbb; hq[2 * kNH + o + j] = sab;
      }
    }
    __syncthreads();
    if (vrows > 0) {
      const int o0 = vy0 * kTW + vx;
      uint32_t sa = 0, sb = 0;
      unsigned long long saa = 0, sbb = 0, sab = 0;
#pragma unroll
      for (int dy = 0; dy < kWin; ++dy) {
        const int o = o0 + dy * kTW;
        sa += hb[o]; sb += hb[kNH + o]; saa += hq[o]; sbb += hq[kNH + o]; sab += hq[2 * kNH + o];
      }
      const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
      const double inv49k = 1.0 / (49.0 * (double)kK);
      const double invv = 1.0 / (48.0 * 49.0 * (double)kK * (double)kK);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < vrows) {
          if (j > 0) {
            const int on = o0 + (j + kWin - 1) * kTW, oo = o0 + (j - 1) * kTW;
            sa += hb[on] - hb[oo]; sb += hb[kNH + on] - hb[kNH + oo];
            saa += hq[on] - hq[oo]; sbb += hq[kNH + on] - hq[kNH + oo]; sab += hq[2 * kNH + on] - hq[2 * kNH + oo];
          }
          if (col_ok && wy0 + vy0 + j < nwy) {
            // 49 Sab - Sa Sb etc. are exact in int64 (49 * 49 * Z^2 < 7.5e18): the variances never see a rounded sum
            const unsigned long long pa = (unsigned long long)sa * sa, pb = (unsigned long long)sb * sb;
            const long long Txy = (long long)(49ull * sab) - (long long)((unsigned long long)sa * sb);
            const long long Txx = (long long)(49ull * saa - pa), Tyy = (long long)(49ull * sbb - pb);
            const double ux = 16.0 + (double)sa * inv49k, uy = 16.0 + (double)sb * inv49k;
            const double A1 = 2.0 * ux * uy + C1, B1 = ux * ux + uy * uy + C1;
            const double A2 = 2.0 * ((double)Txy * invv) + C2;
            const double B2 = ((double)Txx + (double)Tyy) * invv + C2;
            acc_y += __double2ll_rn(fast_div(A1 * A2, B1 * B2) * kFix);
          }
        }
      }
    }
  }

  // ---- block reduction (integers: order does not matter), integer atomics, last block converts
  long long v5[5] = {acc_y, acc_r, acc_g, acc_b, (long long)ssd};
#pragma unroll
  for (int k = 0; k < 5; ++k)
    for (int o = 16; o > 0; o >>= 1) v5[k] += __shfl_xor_sync(0xffffffffu, v5[k], o);
  if ((tid & 31) == 0)
    for (int k = 0; k < 5; ++k) red[tid >> 5][k] = v5[k];
  __syncthreads();
  if (tid == 0) {
    unsigned long long t[5];
    for (int k = 0; k < 5; ++k) {
      long long s_ = 0;
      for (int wv = 0; wv < 8; ++wv) s_ += red[wv][k];
      t[k] = (unsigned long long)s_;
    }
    unsigned long long* gacc = reinterpret_cast<unsigned long long*>(res->acc);
    for (int k = 0; k < 4; ++k) atomicAdd(gacc + k, t[k]);
    atomicAdd(gacc + 4, t[4] & 0xffffffffull);     // squared error: low / high halves (the total exceeds 2^64)
    atomicAdd(gacc + 5, t[4] >> 32);
    __threadfence();
    const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(&res->ticket), 1ull);
    if (done == (unsigned long long)gridDim.x * gridDim.y - 1) {
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

extern "C" int sr_score_pair_u8(const uint8_t* a, const uint8_t* b, int h, int w, int crop,
                                sr_score_result* result, void* stream) {
  if (!a || !b || !result) return set_error(SR_ERR_INVALID, "sr_score_pair_u8: null pointer");
  if (crop < 0 || h - 2 * crop < 7 || w - 2 * crop < 7)
    return set_error(SR_ERR_INVALID, "sr_score_pair_u8: image smaller than the 7x7 SSIM window after cropping");
  const int ch_ = h - 2 * crop, cw_ = w - 2 * crop;
  dim3 grid((cw_ + kTW - 1) / kTW, (ch_ + kTH - 1) / kTH);
  if (grid.y > 65535u) return set_error(SR_ERR_UNSUPPORTED, "sr_score_pair_u8: image taller than 1.7 M rows");
  static unsigned long long attr_done = 0;
  if (int rc = ensure_dynamic_smem(score_pair_kernel, kScoreSmem, &attr_done, "cudaFuncSetAttribute(score_pair_kernel)"))
    return rc;
  score_pair_kernel<<<grid, 256, kScoreSmem, as_stream(stream)>>>(a, b, h, w, crop, result);
  return check_launch("score_pair_kernel");
}
