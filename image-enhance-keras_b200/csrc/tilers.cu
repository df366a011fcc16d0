// The reference's two alternative tilers (BaseSuperResolutionModel.upscalePatch models.py:419-604 and
// upscale(mode='patch') models.py:645-680, 758-790) around the conv stack:
//   * every selected p x p patch of a uint8 image is shrunk x4 with scipy.misc.imresize(..., interp='bicubic')
//     (= optional bytescale contrast stretch of float input + Pillow's 8-bit two-pass bicubic, models.py:490, 672),
//   * the network's p x p outputs are averaged back into the image: img_utils.reconstruct_from_patches_2dlocal
//     (img_utils.py:442-511: interior patches contribute their [4, p-4) window only, division by the count map) or
//     sklearn's reconstruct_from_patches_2d behind img_utils.combine_patches (img_utils.py:189-193: all dense
//     patches, division by the closed-form overlap count).
// Both are integer / fixed-order arithmetic and are reproduced bit for bit: the bicubic pass uses Pillow's
// fixed-point coefficients (precomputed on the host exactly as Resample.c does, 22 fractional bits, uint8
// intermediate between the horizontal and the vertical pass); the averaging sums float64 in the reference's own
// (i, j) lexicographic patch order.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "internal.h"

namespace sr {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)min(max(v, 0), 255);
}

// One block per patch.  Patch n = a * cnt_w + b sits at (a * step, b * step) of the uint8 image [H][W][3].
// smem: patch [p][p*3] u8 | horizontal result [p][q*3] u8 | reduction scratch.
__global__ void __launch_bounds__(256)
patch_down4_kernel(const uint8_t* __restrict__ img, int H, int W, int p, int step, int cnt_h, int cnt_w,
                   long long n0, long long n1, int stretch, const int* __restrict__ bounds,
                   const int* __restrict__ kk, int ksize, float divisor, float* __restrict__ out) {
  extern __shared__ uint8_t sm[];
  const int q = p >> 2, row = p * 3, qrow = q * 3;
  uint8_t* patch = sm;
  uint8_t* hres = sm + (size_t)p * row;
  __shared__ int s_min[8], s_max[8];
  for (long long n = n0 + blockIdx.x; n < n1; n += gridDim.x) {
    const int a = (int)(n / cnt_w), b = (int)(n - (long long)a * cnt_w);
    const uint8_t* src = img + ((size_t)a * step * W + (size_t)b * step) * 3;
    int lo = 255, hi = 0;
    for (int i = threadIdx.x; i < p * row; i += blockDim.x) {
      const int y = i / row, e = i - y * row;
      const uint8_t v = src[(size_t)y * W * 3 + e];
      patch[i] = v;
      lo = min(lo, (int)v);
      hi = max(hi, (int)v);
    }
    if (stretch) {  // scipy.misc.bytescale of the float64 patch: (v - cmin) * (255 / (cmax - cmin)), clip, + 0.5, trunc
      for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
      if ((threadIdx.x & 31) == 0) {
        s_min[threadIdx.x >> 5] = lo;
        s_max[threadIdx.x >> 5] = hi;
      }
      __syncthreads();
      lo = 255, hi = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        lo = min(lo, s_min[w]);
        hi = max(hi, s_max[w]);
      }
      const int cs = hi - lo;
      const double scale = 255.0 / (double)(cs == 0 ? 1 : cs);
      for (int i = threadIdx.x; i < p * row; i += blockDim.x) {
        double v = (double)((int)patch[i] - lo) * scale + 0.0;
        v = fmin(fmax(v, 0.0), 255.0) + 0.5;
        patch[i] = (uint8_t)v;
      }
    }
    __syncthreads();
    // horizontal pass: hres[y][xx][c] = clip8(2^21 + sum_x patch[y][xmin + x][c] * k[xx][x])
    for (int i = threadIdx.x; i < p * qrow; i += blockDim.x) {
      const int y = i / qrow, r = i - y * qrow;
      const int xx = r / 3, c = r - xx * 3;
      const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
      const int* k = kk + xx * ksize;
      int acc = 1 << (kPrecisionBits - 1);
      const uint8_t* s = patch + y * row + xmin * 3 + c;
      for (int x = 0; x < cnt; ++x) acc += (int)s[x * 3] * k[x];
      hres[i] = clip8(acc);
    }
    __syncthreads();
    // vertical pass, then / divisor
    float* o = out + (size_t)(n - n0) * q * qrow;
    for (int i = threadIdx.x; i < q * qrow; i += blockDim.x) {
      const int yy = i / qrow, r = i - yy * qrow;
      const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
      const int* k = kk + yy * ksize;
      int acc = 1 << (kPrecisionBits - 1);
      const uint8_t* s = hres + ymin * qrow + r;
      for (int y = 0; y < cnt; ++y) acc += (int)s[y * qrow] * k[y];
      o[i] = __fdiv_rn((float)clip8(acc), divisor);
    }
    __syncthreads();
  }
}

// sum[Y][X][c] += sum over the patches (a, b), a in [a0, a1), in (a, b) lexicographic order, that cover (Y, X) of
// (double)(float)(patches[(a - a0) * cnt_w + b][Y - a*step][X - b*step][c] * mul).  A patch is "interior" when
// a > 0, b > 0, a != edge_a and b != edge_b (edge = the grid index that sits on the last dense position n - 1 of the
// reference's `i < n_h - 1` test, or -1 when the stepped grid does not reach it); interior patches contribute only
// rows / columns [pad, P - pad).
// count[Y][X] += number of contributions (written from channel 0).  One thread per (Y, X, c).
__global__ void patch_average_accumulate_kernel(const float* __restrict__ patches, int P, int step, int pad,
                                                int cnt_h, int cnt_w, int a0, int a1, int edge_a, int edge_b, float mul,
                                                int out_h, int out_w, double* __restrict__ sum,
                                                int* __restrict__ count) {
  // only the image rows this band of patches can touch: [a0*step, (a1-1)*step + P)
  const int y_first = a0 * step, y_last = min(out_h, (a1 - 1) * step + P);
  const size_t total = (size_t)(y_last - y_first) * out_w * 3;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t idx = t + (size_t)y_first * out_w * 3;
    const int c = (int)(idx % 3);
    const size_t px = idx / 3;
    const int X = (int)(px % out_w), Y = (int)(px / out_w);
    // patches whose full window covers Y: a*step <= Y < a*step + P
    int alo = Y - P + 1 <= 0 ? 0 : (Y - P + step) / step;
    int ahi = min(Y / step, cnt_h - 1);
    int blo = X - P + 1 <= 0 ? 0 : (X - P + step) / step;
    int bhi = min(X / step, cnt_w - 1);
    alo = max(alo, a0);
    ahi = min(ahi, a1 - 1);
    double acc = sum[idx];  // continue the running float64 sum: the additions keep the reference's order across calls
    int n = 0;
    for (int a = alo; a <= ahi; ++a) {
      const int dy = Y - a * step;
      const bool a_in = a > 0 && a != edge_a;
      for (int b = blo; b <= bhi; ++b) {
        const int dx = X - b * step;
        if (pad > 0 && a_in && b > 0 && b != edge_b && (dy < pad || dy >= P - pad || dx < pad || dx >= P - pad))
          continue;
        const float v = patches[((((size_t)(a - a0) * cnt_w + b) * P + dy) * P + dx) * 3 + c];
        acc += (double)__fmul_rn(v, mul);
        ++n;
      }
    }
    sum[idx] = acc;
    if (c == 0) count[px] += n;
  }
}

// img = sum / count (float64, as numpy divides), u8 = np.clip(img, 0, 255).astype('uint8') (truncation).
// closed_P > 0: divide by sklearn's closed-form overlap count min(Y+1, P, H-Y) * min(X+1, P, W-X)
// (reconstruct_from_patches_2d; it exceeds the true count on images smaller than 2P-1, and so does the reference).
__global__ void patch_average_finalize_kernel(const double* __restrict__ sum, const int* __restrict__ count,
                                              int out_h, int out_w, int closed_P, double* __restrict__ out_f64,
                                              uint8_t* __restrict__ out_u8) {
  const size_t npix = (size_t)out_h * out_w;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < npix * 3;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t px = idx / 3;
    double den;
    if (closed_P > 0) {
      const int X = (int)(px % out_w), Y = (int)(px / out_w);
      den = (double)(min(min(Y + 1, closed_P), out_h - Y) * min(min(X + 1, closed_P), out_w - X));
    } else {
      den = (double)count[px];
    }
    const double v = sum[idx] / den;
    if (out_f64) out_f64[idx] = v;
    if (out_u8) out_u8[idx] = (uint8_t)fmin(fmax(v, 0.0), 255.0);  // NaN (uncovered pixel) -> 0
  }
}

}  // namespace
}  // namespace sr

using namespace sr;

extern "C" int sr_patch_down4_u8(const uint8_t* img, int H, int W, int p, int step, int cnt_h, int cnt_w,
                                 long long n0, long long n1, int stretch, const int* bounds, const int* kk,
                                 int ksize, float divisor, float* out_f32, void* stream) {
  if (!img || !bounds || !kk || !out_f32) return set_error(SR_ERR_INVALID, "sr_patch_down4_u8: null pointer");
  if (p < 4 || (p & 3) || p > 160) return set_error(SR_ERR_UNSUPPORTED, "sr_patch_down4_u8: patch size must be a multiple of 4 in [4,160]");
  if (step < 1 || cnt_h < 1 || cnt_w < 1 || (cnt_h - 1) * step + p > H || (cnt_w - 1) * step + p > W)
    return set_error(SR_ERR_INVALID, "sr_patch_down4_u8: the patch grid does not fit the image");
  if (n0 < 0 || n1 > (long long)cnt_h * cnt_w || n0 > n1) return set_error(SR_ERR_INVALID, "sr_patch_down4_u8: bad patch range");
  if (n0 == n1) return SR_OK;
  const size_t smem = (size_t)p * p * 3 + (size_t)p * (p / 4) * 3;
  static size_t attr_set = 0;
  if (smem > 48 * 1024 && smem > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(patch_down4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(patch_down4_kernel)");
    attr_set = smem;
  }
  const long long n = n1 - n0;
  const unsigned grid = (unsigned)(n < 148 * 16 ? n : 148 * 16);
  patch_down4_kernel<<<grid, 256, smem, as_stream(stream)>>>(img, H, W, p, step, cnt_h, cnt_w, n0, n1, stretch, bounds,
                                                            kk, ksize, divisor, out_f32);
  return check_launch("patch_down4_kernel");
}

extern "C" int sr_patch_average_accumulate(const float* patches, int P, int step, int pad, int cnt_h, int cnt_w,
                                           int a0, int a1, int edge_a, int edge_b, float mul, int out_h, int out_w,
                                           double* sum, int* count, void* stream) {
  if (!patches || !sum || !count) return set_error(SR_ERR_INVALID, "sr_patch_average_accumulate: null pointer");
  if (P < 1 || step < 1 || pad < 0 || 2 * pad > P || cnt_h < 1 || cnt_w < 1 || a0 < 0 || a1 > cnt_h || a0 > a1)
    return set_error(SR_ERR_INVALID, "sr_patch_average_accumulate: bad geometry");
  if (out_h < (cnt_h - 1) * step + P || out_w < (cnt_w - 1) * step + P)
    return set_error(SR_ERR_INVALID, "sr_patch_average_accumulate: the patch grid does not fit the image");
  if (a0 == a1) return SR_OK;
  const int y_last = std::min(out_h, (a1 - 1) * step + P);
  const size_t total = (size_t)(y_last - a0 * step) * out_w * 3;
  patch_average_accumulate_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, as_stream(stream)>>>(
      patches, P, step, pad, cnt_h, cnt_w, a0, a1, edge_a, edge_b, mul, out_h, out_w, sum, count);
  return check_launch("patch_average_accumulate_kernel");
}

extern "C" int sr_patch_average_finalize(const double* sum, const int* count, int out_h, int out_w, int closed_P,
                                         double* out_f64, uint8_t* out_u8, void* stream) {
  if (!sum || !count || (!out_f64 && !out_u8)) return set_error(SR_ERR_INVALID, "sr_patch_average_finalize: null pointer");
  if (out_h < 1 || out_w < 1 || closed_P < 0) return set_error(SR_ERR_INVALID, "sr_patch_average_finalize: bad size");
  const size_t npix = (size_t)out_h * out_w;
  patch_average_finalize_kernel<<<grid_for(npix * 3, 256, 148 * 16), 256, 0, as_stream(stream)>>>(
      sum, count, out_h, out_w, closed_P, out_f64, out_u8);
  return check_launch("patch_average_finalize_kernel");
}
