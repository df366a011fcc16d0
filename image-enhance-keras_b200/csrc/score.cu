// Y-channel PSNR + SSIM scoring (reference: scorpath.py:174-228, PSNR.py:54-84, skimage
// rgb2ycbcr / compare_ssim semantics restated in oracle/scoring.py).
// One fused pass: uint8 RGB pair -> crop -> {Y,R,G,B} planes in shared memory (fp64) -> per-window
// 7x7 moment sums -> SSIM map value -> block reduction (warp shuffles) -> fp64 atomics.
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.h"

namespace sr {
namespace {

constexpr int kTile = 32;          // window centres per block edge
constexpr int kWin = 7;
constexpr int kHalo = kTile + kWin - 1;  // 38

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

__global__ void __launch_bounds__(256) score_pair_kernel(const uint8_t* __restrict__ a,
                                                         const uint8_t* __restrict__ b, int h, int w,
                                                         int crop, sr_score_result* __restrict__ res) {
  __shared__ double pa[kHalo * kHalo];
  __shared__ double pb[kHalo * kHalo];
  __shared__ double red[32];
  __shared__ double lut[256];   // v / 255.0 for every uint8 v: the correctly rounded quotients img_as_float produces
  lut[threadIdx.x] = (double)threadIdx.x / 255.0;
  const int ch_ = h - 2 * crop, cw_ = w - 2 * crop;  // cropped size
  const int wy0 = blockIdx.y * kTile, wx0 = blockIdx.x * kTile;  // first window (top-left) of the block
  const int nwy = ch_ - kWin + 1, nwx = cw_ - kWin + 1;          // number of valid windows per axis
  const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
  const double cov_norm = 49.0 / 48.0;

  double sum_sq = 0.0;
  double ssim_acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int plane = 0; plane < 4; ++plane) {  // 0: Y, 1..3: R,G,B
    __syncthreads();
    for (int i = threadIdx.x; i < kHalo * kHalo; i += blockDim.x) {
      const int ly = i / kHalo, lx = i - ly * kHalo;
      const int y = wy0 + ly, x = wx0 + lx;  // cropped coordinates
      double va = 0.0, vb = 0.0;
      if (y < ch_ && x < cw_) {
        const size_t o = ((size_t)(y + crop) * w + (x + crop)) * 3;
        if (plane == 0) {
          // same operations in the same order as y_from_rgb, the three divisions looked up
          va = (lut[a[o]] * 65.481 + lut[a[o + 1]] * 128.553 + lut[a[o + 2]] * 24.966) + 16.0;
          vb = (lut[b[o]] * 65.481 + lut[b[o + 1]] * 128.553 + lut[b[o + 2]] * 24.966) + 16.0;
          // squared error: every cropped pixel is owned by exactly one block (its tile interior)
          if (ly < kTile && lx < kTile) {
            const double d = va - vb;
            sum_sq += d * d;
          }
        } else {
          va = (double)a[o + plane - 1];
          vb = (double)b[o + plane - 1];
        }
      }
      if (plane == 0) {
        pa[i] = va;
        pb[i] = vb;
      } else {          // R, G, B hold integers 0..255: fp32 keeps every window sum (< 2^24) exact
        reinterpret_cast<float*>(pa)[i] = (float)va;
        reinterpret_cast<float*>(pb)[i] = (float)vb;
      }
    }
    __syncthreads();
    // Separable window sums: a thread owns window column lx and 4 consecutive window rows.  For each of the 10
    // halo rows under them it forms the horizontal 7-sums of (x, y, x^2, y^2, xy) once and adds them to the windows
    // that contain the row (2.8x fewer shared-memory reads than summing 49 taps per window; R/G/B planes hold
    // integers, so their sums are exact in any order).
    {
      const int lx = threadIdx.x & 31, ly0 = (threadIdx.x >> 5) * 4;
      double acc[4][5];
#pragma unroll
      for (int wv = 0; wv < 4; ++wv)
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[wv][k] = 0.0;
      if (plane == 0) {
#pragma unroll
        for (int r = 0; r < 4 + kWin - 1; ++r) {
          const double* ra = pa + (ly0 + r) * kHalo + lx;
          const double* rb = pb + (ly0 + r) * kHalo + lx;
          double hx = 0, hy = 0, hxx = 0, hyy = 0, hxy = 0;
#pragma unroll
          for (int dx = 0; dx < kWin; ++dx) {
            const double x = ra[dx], y = rb[dx];
            hx += x; hy += y; hxx += x * x; hyy += y * y; hxy += x * y;
          }
#pragma unroll
          for (int wv = 0; wv < 4; ++wv) {
            if (r >= wv && r < wv + kWin) {
              acc[wv][0] += hx; acc[wv][1] += hy; acc[wv][2] += hxx; acc[wv][3] += hyy; acc[wv][4] += hxy;
            }
          }
        }
      } else {
        float facc[4][5];
#pragma unroll
        for (int wv = 0; wv < 4; ++wv)
#pragma unroll
          for (int k = 0; k < 5; ++k) facc[wv][k] = 0.f;
#pragma unroll
        for (int r = 0; r < 4 + kWin - 1; ++r) {
          const float* ra = reinterpret_cast<const float*>(pa) + (ly0 + r) * kHalo + lx;
          const float* rb = reinterpret_cast<const float*>(pb) + (ly0 + r) * kHalo + lx;
          float hx = 0, hy = 0, hxx = 0, hyy = 0, hxy = 0;
#pragma unroll
          for (int dx = 0; dx < kWin; ++dx) {
            const float x = ra[dx], y = rb[dx];
            hx += x; hy += y; hxx = fmaf(x, x, hxx); hyy = fmaf(y, y, hyy); hxy = fmaf(x, y, hxy);
          }
#pragma unroll
          for (int wv = 0; wv < 4; ++wv) {
            if (r >= wv && r < wv + kWin) {
              facc[wv][0] += hx; facc[wv][1] += hy; facc[wv][2] += hxx; facc[wv][3] += hyy; facc[wv][4] += hxy;
            }
          }
        }
#pragma unroll
        for (int wv = 0; wv < 4; ++wv)
#pragma unroll
          for (int k = 0; k < 5; ++k) acc[wv][k] = (double)facc[wv][k];
      }
#pragma unroll
      for (int wv = 0; wv < 4; ++wv) {
        if (wy0 + ly0 + wv >= nwy || wx0 + lx >= nwx) continue;
        // means by multiplication (fp64 division is ~20 instructions; 1/49 rounded once costs 1e-16 relative)
        constexpr double inv49 = 1.0 / 49.0;
        const double ux = acc[wv][0] * inv49, uy = acc[wv][1] * inv49;
        const double uxx = acc[wv][2] * inv49, uyy = acc[wv][3] * inv49, uxy = acc[wv][4] * inv49;
        const double vx = cov_norm * (uxx - ux * ux);
        const double vy = cov_norm * (uyy - uy * uy);
        const double vxy = cov_norm * (uxy - ux * uy);
        const double A1 = 2 * ux * uy + C1, A2 = 2 * vxy + C2;
        const double B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
        ssim_acc[plane] += (A1 * A2) / (B1 * B2);
      }
    }
  }
  const double t_sq = block_sum(sum_sq, red);
  const double t0 = block_sum(ssim_acc[0], red);
  const double t1 = block_sum(ssim_acc[1], red);
  const double t2 = block_sum(ssim_acc[2], red);
  const double t3 = block_sum(ssim_acc[3], red);
  if (threadIdx.x == 0) {
    atomicAdd(&res->sum_sq_y, t_sq);
    atomicAdd(&res->ssim_y_sum, t0);
    atomicAdd(&res->ssim_rgb_sum[0], t1);
    atomicAdd(&res->ssim_rgb_sum[1], t2);
    atomicAdd(&res->ssim_rgb_sum[2], t3);
    if (blockIdx.x == 0 && blockIdx.y == 0) {
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
  dim3 grid((cw_ + kTile - 1) / kTile, (ch_ + kTile - 1) / kTile);
  score_pair_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, b, h, w, crop, result);
  return check_launch("score_pair_kernel");
}
