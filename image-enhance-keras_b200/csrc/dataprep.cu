// Dataset preparation on the device: the kernels behind img_utils.transform_images (reference img_utils.py:44-123).
// Byte / integer / float64 work restated from the libraries the reference calls (Pillow Resample.c and Filter.c,
// scipy.ndimage correlate1d, scipy.misc.bytescale); every kernel is bit-exact against them (oracle/dataprep.py,
// pinned to the installed Pillow / scipy and to the reference function's own output).
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.h"

namespace sr {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow's fixed point for 8-bit resampling

// One pass of Pillow's 8-bpc resize (ImagingResampleHorizontal_8bpc / Vertical_8bpc) over NB images:
// out = clip8((2^21 + sum_t src[first + t] * kk[o][t]) >> 22), int32 arithmetic like the C code.
// AXIS 1: src [NB,H,W,3] -> dst [NB,H,out_n,3]; AXIS 0: src [NB,H,W,3] -> dst [NB,out_n,W,3].
template <int AXIS>
__global__ void __launch_bounds__(256)
resample_pass_u8_kernel(const uint8_t* __restrict__ src, int NB, int H, int W, int out_n,
                        const int* __restrict__ bounds, const int* __restrict__ kk, int ksize,
                        uint8_t* __restrict__ dst) {
  const int OH = AXIS == 0 ? out_n : H, OW = AXIS == 1 ? out_n : W;
  const size_t total = (size_t)NB * OH * OW * 3;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const unsigned per_img = (unsigned)OH * OW * 3;
    const unsigned n = (unsigned)(idx / per_img), r = (unsigned)(idx - (size_t)n * per_img);
    const unsigned c = r % 3u, px = r / 3u;
    const unsigned x = px % (unsigned)OW, y = px / (unsigned)OW;
    const int o = AXIS == 1 ? (int)x : (int)y;
    const int first = bounds[2 * o], cnt = bounds[2 * o + 1];
    const int* k = kk + (size_t)o * ksize;
    const uint8_t* s = src + (size_t)n * H * W * 3 + c;
    int acc = 1 << (kPrecisionBits - 1);
    if (AXIS == 1) {
      s += ((size_t)y * W + first) * 3;
      for (int t = 0; t < cnt; ++t) acc += (int)s[(size_t)t * 3] * k[t];
    } else {
      s += ((size_t)first * W + x) * 3;
      for (int t = 0; t < cnt; ++t) acc += (int)s[(size_t)t * W * 3] * k[t];
    }
    acc >>= kPrecisionBits;
    dst[idx] = (uint8_t)(acc < 0 ? 0 : acc > 255 ? 255 : acc);
  }
}

// PIL ImageFilter.SHARPEN (Filter.c ImagingFilter3x3, kernel (-2 x8, 32)/16, offset 0): interior pixels
// clip8(0.5 + 2c - S8/8) with truncation -- all terms are multiples of 1/8 below 2^11, exact in the C code's
// float32 in any order, so integers give the same result; the one-pixel border is copied.
__global__ void __launch_bounds__(256)
sharpen3x3_u8_kernel(const uint8_t* __restrict__ src, int NB, int H, int W, uint8_t* __restrict__ dst) {
  const size_t total = (size_t)NB * H * W * 3;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const unsigned per_img = (unsigned)H * W * 3;
    const unsigned r = (unsigned)(idx % per_img);
    const unsigned px = r / 3u;
    const int x = (int)(px % (unsigned)W), y = (int)(px / (unsigned)W);
    const int c0 = src[idx];
    if (x == 0 || y == 0 || x == W - 1 || y == H - 1) {
      dst[idx] = (uint8_t)c0;
      continue;
    }
    const uint8_t* p = src + idx;
    const int rs = W * 3;
    const int s8 = p[-rs - 3] + p[-rs] + p[-rs + 3] + p[-3] + p[3] + p[rs - 3] + p[rs] + p[rs + 3];
    const int v8 = 16 * c0 - s8 + 4;                    // 8 * (0.5 + 2c - S8/8)
    dst[idx] = (uint8_t)(v8 <= 0 ? 0 : min(v8 >> 3, 255));
  }
}

__device__ __forceinline__ void block_minmax(double& lo, double& hi, double* red) {
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { red[2 * wid] = lo; red[2 * wid + 1] = hi; }
  __syncthreads();
  lo = red[0]; hi = red[1];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { lo = fmin(lo, red[2 * i]); hi = fmax(hi, red[2 * i + 1]); }
}

// scipy.misc.bytescale of a float64 value: ((v - cmin) * (255 / (cmax - cmin))).clip(0, 255) + 0.5 -> uint8
__device__ __forceinline__ uint8_t bytescale_f64(double v, double cmin, double scale) {
  double b = __dmul_rn(__dsub_rn(v, cmin), scale);
  b = b < 0.0 ? 0.0 : b > 255.0 ? 255.0 : b;
  return (uint8_t)(int)__dadd_rn(b, 0.5);
}

// reflect index (d c b a | a b c d | d c b a) for -n <= i < 2n
__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -1 - i : i >= n ? 2 * n - 1 - i : i; }

// One block per sub-image (img_utils.py:96-117): ip = float64(img[x:x+P, y:y+P, :]);
//   y_u8 = bytescale(ip)                              (imsave of the float64 sample)
//   g_u8 = bytescale(gaussian_filter(ip, sigma))      (the toimage step of the imresize that follows)
// The Gaussian is scipy.ndimage's: correlate1d along axes 0, 1, 2 (the channel axis included), radius R, reflect,
// symmetric-kernel order tmp = x0*w0; tmp += (x[-j] + x[j]) * w[j] for j = R..1, float64, no contraction.
__global__ void __launch_bounds__(256)
dataprep_patch_kernel(const uint8_t* __restrict__ img, int H, int W, const int* __restrict__ pos, int P,
                      const double* __restrict__ wts, int R, uint8_t* __restrict__ y_u8,
                      uint8_t* __restrict__ g_u8) {
  extern __shared__ double dp_smem[];
  __shared__ double red[64];
  const int n3 = P * P * 3;
  double* a = dp_smem;
  double* b = dp_smem + n3;
  const int px = pos[2 * blockIdx.x], py = pos[2 * blockIdx.x + 1];   // first row / first column
  double lo = 1e300, hi = -1e300;
  for (int i = threadIdx.x; i < n3; i += blockDim.x) {
    const int c = i % 3, j = (i / 3) % P, r = i / (3 * P);
    const double v = (double)img[((size_t)(px + r) * W + (py + j)) * 3 + c];
    a[i] = v;
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
  block_minmax(lo, hi, red);
  {
    double cs = __dsub_rn(hi, lo);
    if (cs == 0.0) cs = 1.0;
    const double scale = __ddiv_rn(255.0, cs);
    uint8_t* dst = y_u8 + (size_t)blockIdx.x * n3;
    for (int i = threadIdx.x; i < n3; i += blockDim.x) dst[i] = bytescale_f64(a[i], lo, scale);
  }
  // three correlate1d passes: a -> b (rows), b -> a (columns), a -> b (channels)
  const int strides[3] = {3 * P, 3, 1}, lens[3] = {P, P, 3};
  double* src = a;
  double* dstb = b;
  for (int axis = 0; axis < 3; ++axis) {
    __syncthreads();
    const int st = strides[axis], len = lens[axis];
    for (int i = threadIdx.x; i < n3; i += blockDim.x) {
      const int k = (i / st) % len;              // coordinate along the axis
      const double* line = src + (i - k * st);   // element 0 of this line
      double tmp = __dmul_rn(line[k * st], wts[0]);
      for (int j = R; j >= 1; --j) {
        const double s = __dadd_rn(line[reflect(k - j, len) * st], line[reflect(k + j, len) * st]);
        tmp = __dadd_rn(tmp, __dmul_rn(s, wts[j]));
      }
      dstb[i] = tmp;
    }
    double* t = src; src = dstb; dstb = t;
  }
  __syncthreads();
  // src now holds the filtered sample
  lo = 1e300; hi = -1e300;
  for (int i = threadIdx.x; i < n3; i += blockDim.x) { lo = fmin(lo, src[i]); hi = fmax(hi, src[i]); }
  block_minmax(lo, hi, red);
  double cs = __dsub_rn(hi, lo);
  if (cs == 0.0) cs = 1.0;
  const double scale = __ddiv_rn(255.0, cs);
  uint8_t* dst = g_u8 + (size_t)blockIdx.x * n3;
  for (int i = threadIdx.x; i < n3; i += blockDim.x) dst[i] = bytescale_f64(src[i], lo, scale);
}

}  // namespace
}  // namespace sr

using namespace sr;

extern "C" int sr_resize_u8(const uint8_t* src, int NB, int H, int W, int out_h, int out_w,
                            const int* bounds_x, const int* kk_x, int ksize_x, const int* bounds_y,
                            const int* kk_y, int ksize_y, uint8_t* tmp, uint8_t* dst, void* stream) {
  if (!src || !dst) return set_error(SR_ERR_INVALID, "sr_resize_u8: null pointer");
  if (NB < 1 || H < 1 || W < 1 || out_h < 1 || out_w < 1) return set_error(SR_ERR_INVALID, "sr_resize_u8: bad size");
  const bool need_x = out_w != W, need_y = out_h != H;
  if ((need_x && (!bounds_x || !kk_x || ksize_x < 1)) || (need_y && (!bounds_y || !kk_y || ksize_y < 1)))
    return set_error(SR_ERR_INVALID, "sr_resize_u8: missing coefficient table");
  if (need_x && need_y && !tmp) return set_error(SR_ERR_INVALID, "sr_resize_u8: tmp [NB,H,out_w,3] needed for two passes");
  cudaStream_t st = as_stream(stream);
  if (!need_x && !need_y) {
    cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)NB * H * W * 3, cudaMemcpyDeviceToDevice, st);
    return e == cudaSuccess ? SR_OK : set_cuda_error(e, "sr_resize_u8 copy");
  }
  const uint8_t* cur = src;
  if (need_x) {   // Pillow: horizontal pass first, uint8 intermediate
    uint8_t* o = need_y ? tmp : dst;
    resample_pass_u8_kernel<1><<<grid_for((size_t)NB * H * out_w * 3, 256, 148 * 16), 256, 0, st>>>(
        cur, NB, H, W, out_w, bounds_x, kk_x, ksize_x, o);
    const int rc = check_launch("resample_pass_u8_kernel<horizontal>");
    if (rc != SR_OK) return rc;
    cur = o;
  }
  if (need_y) {
    const int Wc = need_x ? out_w : W;
    resample_pass_u8_kernel<0><<<grid_for((size_t)NB * out_h * Wc * 3, 256, 148 * 16), 256, 0, st>>>(
        cur, NB, H, Wc, out_h, bounds_y, kk_y, ksize_y, dst);
    return check_launch("resample_pass_u8_kernel<vertical>");
  }
  return SR_OK;
}

extern "C" int sr_sharpen3x3_u8(const uint8_t* src, int NB, int H, int W, uint8_t* dst, void* stream) {
  if (!src || !dst || src == dst) return set_error(SR_ERR_INVALID, "sr_sharpen3x3_u8: null or aliased pointer");
  if (NB < 1 || H < 1 || W < 1) return set_error(SR_ERR_INVALID, "sr_sharpen3x3_u8: bad size");
  sharpen3x3_u8_kernel<<<grid_for((size_t)NB * H * W * 3, 256, 148 * 16), 256, 0, as_stream(stream)>>>(src, NB, H, W, dst);
  return check_launch("sharpen3x3_u8_kernel");
}

extern "C" int sr_dataprep_patches(const uint8_t* img, int H, int W, const int* pos, int n_patches, int P,
                                   const double* weights, int radius, uint8_t* y_u8, uint8_t* g_u8, void* stream) {
  if (!img || !pos || !weights || !y_u8 || !g_u8) return set_error(SR_ERR_INVALID, "sr_dataprep_patches: null pointer");
  if (n_patches < 1) return SR_OK;
  if (P < 1 || P > H || P > W || radius < 0 || radius > 8) return set_error(SR_ERR_INVALID, "sr_dataprep_patches: bad size");
  const size_t smem = (size_t)2 * P * P * 3 * sizeof(double);
  if (smem > 200 * 1024) return set_error(SR_ERR_UNSUPPORTED, "sr_dataprep_patches: sub-images larger than 64 x 64");
  static unsigned long long attr_done = 0;
  if (int rc = ensure_dynamic_smem(dataprep_patch_kernel, 200 * 1024, &attr_done, "cudaFuncSetAttribute(dataprep_patch_kernel)"))
    return rc;
  dataprep_patch_kernel<<<n_patches, 256, smem, as_stream(stream)>>>(img, H, W, pos, P, weights, radius, y_u8, g_u8);
  return check_launch("dataprep_patch_kernel");
}
