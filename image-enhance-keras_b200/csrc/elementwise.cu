// Bandwidth-bound kernels of the sr100 hot path: first layer (1x1, K=3), TF1-legacy bilinear x4 and
// its adjoint, patch gather / stitch, depth-to-space, weight repack, casts, loss and Adam.
// All are grid-stride, 128-bit vectorised where the layout allows, sized in multiples of 148 SMs.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "internal.h"

namespace sr {

namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// ------------------------------------------------------------------ head 1x1 conv (3 -> 128) + ReLU
// Reference: models.py:1177 Convolution2D(128,(1,1),activation='relu',name='level1').
// 16 threads per pixel, 8 channels each.  The grid stride is a multiple of 16, so a thread keeps the same 8 channels
// for its whole life: their 24 weights + 8 biases sit in registers, and the loop is unrolled over four pixels whose
// 12-byte inputs are requested together (one dependent load per 768 B written left the kernel latency-bound at
// 0.69 of the copy bandwidth; a pure-write kernel can exceed it: torch.fill reaches 7.3 TB/s on the same box).
__global__ void __launch_bounds__(256)
head1x1_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias, size_t npix,
               uint4* __restrict__ out_bf16, float4* __restrict__ out_f32) {
  const int c0 = (threadIdx.x & 15) * 8;
  float wr[8], wg[8], wb[8], bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    wr[j] = w[c0 + j], wg[j] = w[128 + c0 + j], wb[j] = w[256 + c0 + j];
    bs[j] = bias ? bias[c0 + j] : 0.f;
  }
  const size_t total = npix * 16;  // 16 groups of 8 channels per pixel
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (size_t idx0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx0 < total; idx0 += U * stride) {
    float r[U], g[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t idx = idx0 + u * stride;
      r[u] = g[u] = b[u] = 0.f;
      if (idx < total) {
        const size_t pix = idx >> 4;
        r[u] = in[pix * 3 + 0], g[u] = in[pix * 3 + 1], b[u] = in[pix * 3 + 2];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t idx = idx0 + u * stride;
      if (idx >= total) break;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // accumulate in channel order like a 3-term dot product, then bias, then ReLU
        float a = __fmul_rn(r[u], wr[j]);
        a = __fmaf_rn(g[u], wg[j], a);
        a = __fmaf_rn(b[u], wb[j], a);
        a = __fadd_rn(a, bs[j]);
        o[j] = fmaxf(a, 0.f);
      }
      if (out_bf16)
        out_bf16[idx] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                   pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      if (out_f32) {
        out_f32[idx * 2] = make_float4(o[0], o[1], o[2], o[3]);
        out_f32[idx * 2 + 1] = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
  }
}

// ------------------------------------------------------------------ bilinear x4, TF1 legacy
// Reference: tf.image.resize_bilinear(x, [4h,4w]) (models.py:1392-1399), align_corners=False,
// legacy sampling: src = dst*0.25, lo = floor(src), hi = min(ceil(src), n-1), t = src - lo;
// top = tl + (tr-tl)*tx ; bot = bl + (br-bl)*tx ; out = top + (bot-top)*ty  (fp32, no FMA).
__device__ __forceinline__ float lerp_tf(float a, float b, float t) {
  return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}

template <bool IN_BF16>
__device__ __forceinline__ void load8(const void* base, size_t elem_off, float (&v)[8]) {
  if constexpr (IN_BF16) {
    const uint4 r = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) +
                                                    elem_off);
    v[0] = bf16_lo(r.x); v[1] = bf16_hi(r.x); v[2] = bf16_lo(r.y); v[3] = bf16_hi(r.y);
    v[4] = bf16_lo(r.z); v[5] = bf16_hi(r.z); v[6] = bf16_lo(r.w); v[7] = bf16_hi(r.w);
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
    const float4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}

// One thread = one LR cell (y0,x0) x 8 channels: the four corner values are loaded once and the 4x4 HR
// outputs of the cell are produced from them (top/bot interpolants shared by the four output rows: the
// arithmetic per output is exactly the three-lerp TF formula above).  16 lanes cover the 128 channels of a
// pixel, so every store instruction writes whole 256 B (bf16) / 512 B (fp32) pixels.
template <bool IN_BF16, int R = 4>
__global__ void __launch_bounds__(256)
bilinear4_fwd_kernel(const void* __restrict__ in, const int* __restrict__ src_index, int NB, int H, int W, int C,
                     int CH, int CW, uint4* __restrict__ out_bf16, float4* __restrict__ out_f32) {
  // CH x CW: LR cells actually produced (output extent 4CH x 4CW <= 4H x 4W, the cropped HR stage);
  // src_index: optional gather of the source images (output image n reads input image src_index[n]).
  const int C8 = C >> 3;
  const int OW = R * CW;
  const size_t total = (size_t)NB * CH * CW * C8;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % C8);
    size_t r = idx / C8;
    const int x0 = (int)(r % CW);
    r /= CW;
    const int y0 = (int)(r % CH);
    const int n = (int)(r / CH);
    const int ns = src_index ? src_index[n] : n;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const size_t rowb0 = ((size_t)ns * H + y0) * W, rowb1 = ((size_t)ns * H + y1) * W;
    float tl[8], tr[8], bl[8], br[8];
    load8<IN_BF16>(in, (rowb0 + x0) * C + c8 * 8, tl);
    load8<IN_BF16>(in, (rowb0 + x1) * C + c8 * 8, tr);
    load8<IN_BF16>(in, (rowb1 + x0) * C + c8 * 8, bl);
    load8<IN_BF16>(in, (rowb1 + x1) * C + c8 * 8, br);
    const size_t obase = (((size_t)n * R * CH + R * y0) * OW + R * x0) * C8 + c8;
#pragma unroll
    for (int fx = 0; fx < R; ++fx) {
      const float tx = (float)fx * (1.0f / R);
      float top[8], bot[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        top[j] = lerp_tf(tl[j], tr[j], tx);
        bot[j] = lerp_tf(bl[j], br[j], tx);
      }
#pragma unroll
      for (int fy = 0; fy < R; ++fy) {
        const float ty = (float)fy * (1.0f / R);
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = lerp_tf(top[j], bot[j], ty);
        const size_t oi = obase + ((size_t)fy * OW + fx) * C8;
        if (out_bf16)
          out_bf16[oi] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                    pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
        if (out_f32) {
          out_f32[oi * 2] = make_float4(o[0], o[1], o[2], o[3]);
          out_f32[oi * 2 + 1] = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  }
}

// weight with which HR index Q (0..F*n-1) samples LR index q, along one axis (F = 4 or 2)
template <int F>
__device__ __forceinline__ float axis_weight(int Q, int q, int n) {
  const int lo = Q / F;
  const int fr = Q - lo * F;
  const int hi = fr == 0 ? lo : min(lo + 1, n - 1);
  const float t = (float)fr * (1.f / F);
  float w = 0.f;
  if (lo == q) w += 1.f - t;
  if (hi == q) w += t;
  return w;
}

// Adjoint (gather form, deterministic): gin[y,x] = sum_{Y,X} wy(Y,y) wx(X,x) gout[Y,X] over the (2F-1)^2 HR
// window [Fy-(F-1),Fy+(F-1)] x [Fx-(F-1),Fx+(F-1)] (7x7 for x4, 3x3 for x2).  A warp covers the 128 channels of
// one pixel (512 B per load); the loads of a window row are issued together.
template <int F>
__global__ void __launch_bounds__(256)
bilinear4_bwd_kernel(const float4* __restrict__ gout, int NB, int H, int W, int C,
                     float4* __restrict__ gin) {
  constexpr int WN = 2 * F - 1;
  const int C4 = C >> 2;
  const int OH = F * H, OW = F * W;
  const size_t total = (size_t)NB * H * W * C4;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % C4);
    size_t r = idx / C4;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const int n = (int)(r / H);
    float wx[WN];
    int xi[WN];
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      const int X = F * x - (F - 1) + j;
      const bool ok = X >= 0 && X < OW;
      xi[j] = ok ? X : F * x;
      wx[j] = ok ? axis_weight<F>(X, x, W) : 0.f;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < WN; ++i) {
      const int Y = F * y - (F - 1) + i;
      if (Y < 0 || Y >= OH) continue;
      const float wy = axis_weight<F>(Y, y, H);
      const float4* row = gout + ((size_t)n * OH + Y) * OW * C4 + c4;
      float4 g[WN];
#pragma unroll
      for (int j = 0; j < WN; ++j) g[j] = row[(size_t)xi[j] * C4];
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        const float wgt = wy * wx[j];
        acc.x = fmaf(wgt, g[j].x, acc.x);
        acc.y = fmaf(wgt, g[j].y, acc.y);
        acc.z = fmaf(wgt, g[j].z, acc.z);
        acc.w = fmaf(wgt, g[j].w, acc.w);
      }
    }
    gin[idx] = acc;
  }
}

// ------------------------------------------------------------------ patch gather / stitch
// Reference: img_utils.extract_patches_Step (img_utils.py:601-676): positions
// {x : 0 <= x < dim-p, x % step == 0}, w outer / h inner => n = wi*cnt_h + hi.
template <bool FROM_U8>
__global__ void patch_gather_kernel(const void* __restrict__ src, int h, int w, int canvas_w,
                                    int cnt_h, int cnt_w, int ph, int pw, int step, float divisor,
                                    float* __restrict__ out) {
  const size_t row_elems = (size_t)pw * 3;
  const size_t total = (size_t)cnt_h * cnt_w * ph * row_elems;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx % row_elems);
    size_t r = idx / row_elems;
    const int i = (int)(r % ph);
    const int n = (int)(r / ph);
    const int wi = n / cnt_h, hi = n - wi * cnt_h;
    const int y = hi * step + i;
    const int x = wi * step + e / 3;
    const int c = e % 3;
    float v = 0.f;
    if (FROM_U8) {
      if (y < h && x < w) v = (float)reinterpret_cast<const uint8_t*>(src)[((size_t)y * w + x) * 3 + c];
    } else {
      v = reinterpret_cast<const float*>(src)[((size_t)y * canvas_w + x) * 3 + c];
    }
    out[idx] = divisor == 1.f ? v : __fdiv_rn(v, divisor);
  }
}

// The production gather (uint8 source, 3*pw % 4 == 0): one warp per patch row, any number of same-shaped images per
// launch (image m's patches follow image m-1's).  The row's coordinates are computed once per warp, the lanes
// then cover the row in 16-byte stores (512 B per warp instruction); value / divisor comes from a 256-entry table
// built per block with the same correctly rounded fp32 division, so results are bit-identical to the scalar kernel.
// (The flat-index kernel above -- and its 4-element variant this one replaced -- spends ~250 instructions per
// 16 bytes on 64-bit div/mod and four fp32 divisions: instruction-bound at ~2 TB/s.)
// b / 255 correctly rounded, for a byte b: one product with fl(1/255) and one Newton correction through two FMAs
// (the fast path of a correctly rounded fp32 division); equal to __fdiv_rn((float)b, 255.f) for all 256 inputs
// (tests/test_gpu_tiling.py checks every byte value against numpy's float32 division).
__device__ __forceinline__ float div255(uint32_t b) {
  const float x = (float)b, c = 0x1.010102p-8f;
  const float q = x * c;
  return __fmaf_rn(__fmaf_rn(-q, 255.f, x), c, q);
}

// MODE 0: value / divisor through a 256-entry shared-memory table (any divisor); 1: divisor == 255 in arithmetic
// (no table: 32 random table reads per warp instruction serialise on bank conflicts and made the LSU the
// co-bottleneck of a write-bound kernel); 2: divisor == 1.
template <int MODE>
__global__ void __launch_bounds__(256)
patch_gather_u8_rows_kernel(const uint8_t* __restrict__ imgs, int n_img, size_t img_stride, int h, int w,
                            int cnt_h, int cnt_w, int ph, int pw, int step, float divisor,
                            float4* __restrict__ out) {
  // One block per patch (grid-stride): the patch's coordinates are computed once, its ph * (3*pw/4) 16-byte
  // groups are then walked by all 256 threads with an incrementally updated (row, group) pair -- no division in
  // the loop, every lane busy.  The four bytes of a group come from the one or two aligned 32-bit words that hold
  // them (funnel shift), not from four byte loads.
  __shared__ float lut[MODE == 0 ? 256 : 1];
  if constexpr (MODE == 0) {
    lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, divisor);
    __syncthreads();
  }
  const int row_q = pw * 3 / 4, w3 = w * 3;
  const int patch_q = ph * row_q;
  const unsigned per_img = (unsigned)cnt_h * cnt_w, total = per_img * (unsigned)n_img;
  const int di = 256 / row_q, dq = 256 % row_q;               // +256 groups in (row, group) coordinates
  for (unsigned pidx = blockIdx.x; pidx < total; pidx += gridDim.x) {
    const unsigned m = pidx / per_img, n = pidx - m * per_img;
    const unsigned wi = n / (unsigned)cnt_h, hi = n - wi * (unsigned)cnt_h;
    const int y0 = (int)(hi * step), e_row = (int)(wi * step) * 3;
    const uint8_t* img = imgs + (size_t)m * img_stride;
    float4* dst = out + (size_t)pidx * patch_q;
    int ii = (int)threadIdx.x / row_q, qq = (int)threadIdx.x % row_q;
    constexpr int U = 4;                      // groups per thread and trip: their loads are in flight together
    for (int g = threadIdx.x; g < patch_q; g += 256 * U) {
      uint32_t word[U];
      bool has[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        has[u] = g + 256 * u < patch_q;
        const int y = y0 + ii, e0 = e_row + qq * 4;
        const int nb = (has[u] && y < h) ? min(4, w3 - e0) : 0;          // bytes of the group inside the image row
        word[u] = 0u;
        if (nb > 0) {
          const uintptr_t a = reinterpret_cast<uintptr_t>(img + (size_t)y * w3 + e0);
          const uint32_t sh = (uint32_t)(a & 3u);
          const uint32_t* p = reinterpret_cast<const uint32_t*>(a - sh);
          const uint32_t lo = p[0];
          const uint32_t hi2 = sh + (uint32_t)nb > 4u ? p[1] : 0u;       // second word only when the bytes spill into it
          word[u] = __funnelshift_r(lo, hi2, sh * 8u);
          if (nb < 4) word[u] &= (1u << (8 * nb)) - 1u;
        }
        ii += di;
        qq += dq;
        if (qq >= row_q) { qq -= row_q; ++ii; }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!has[u]) continue;
        const uint32_t b0 = word[u] & 255u, b1 = (word[u] >> 8) & 255u, b2 = (word[u] >> 16) & 255u, b3 = word[u] >> 24;
        float4 v;
        if constexpr (MODE == 0) v = make_float4(lut[b0], lut[b1], lut[b2], lut[b3]);
        else if constexpr (MODE == 1) v = make_float4(div255(b0), div255(b1), div255(b2), div255(b3));
        else v = make_float4((float)b0, (float)b1, (float)b2, (float)b3);
        dst[g + 256 * u] = v;
      }
    }
  }
}

// ------------------------------------------------------------------ minibatch assembly from an HBM-resident dataset
// Reference: img_utils.image_generator (img_utils.py:341-372): batch[i] = imread(file[index[i]]).astype('float32')/255.
// The decoded uint8 images live in HBM ([N][item_bytes]); one launch gathers the rows named by `index` and
// normalises them.  Four bytes per thread: one 4-byte load, one 16-byte store.
template <bool DIV255>
__global__ void __launch_bounds__(256)
batch_gather_u8_kernel(const uint8_t* __restrict__ data, size_t item_bytes,
                       const long long* __restrict__ index, int n, float divisor,
                       float* __restrict__ out) {
  // v / divisor for every uint8 v: a table of the correctly rounded quotients, or (divisor == 255, the reference's
  // case) div255's three fp32 operations -- random table reads serialise on shared-memory bank conflicts
  __shared__ float lut_[DIV255 ? 1 : 256];
  if constexpr (!DIV255) {
    lut_[threadIdx.x] = __fdiv_rn((float)threadIdx.x, divisor);
    __syncthreads();
  }
  auto lut = [&](uint32_t b) { if constexpr (DIV255) return div255(b); else return lut_[b]; };
  const size_t q = item_bytes >> 2;  // uchar4 groups per item
  const size_t total = (size_t)n * q;
  if (total < 0x7fffffffu) {         // 32-bit index arithmetic (64-bit div/mod costs more than the copy)
    // four groups per trip, the loads of each level issued together: index -> bytes -> store is a chain of two
    // dependent loads, and one chain per trip left the kernel latency-bound
    const unsigned q32 = (unsigned)q, tot32 = (unsigned)total, stride = gridDim.x * blockDim.x;
    constexpr int U = 4;
    for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < tot32; i0 += U * stride) {
      unsigned e[U];
      long long ix[U];
      uchar4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned i = i0 + u * stride;
        const unsigned b = i < tot32 ? i / q32 : 0u;     // (i0 + u * stride cannot wrap: tot32 < 2^31)
        e[u] = i - b * q32;
        ix[u] = i < tot32 ? index[b] : -1;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        v[u] = ix[u] >= 0 ? reinterpret_cast<const uchar4*>(data + (size_t)ix[u] * item_bytes)[e[u]] : make_uchar4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (ix[u] >= 0)
          reinterpret_cast<float4*>(out)[i0 + u * stride] = make_float4(lut(v[u].x), lut(v[u].y), lut(v[u].z), lut(v[u].w));
    }
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / q, e = i - b * q;
    const uchar4 v = reinterpret_cast<const uchar4*>(data + (size_t)index[b] * item_bytes)[e];
    reinterpret_cast<float4*>(out)[i] = make_float4(lut(v.x), lut(v.y), lut(v.z), lut(v.w));
  }
}
// Reference: img_utils.rebuild_from_patches_Step (img_utils.py:692-724).  Per axis the owner of
// output coordinate Q is the LAST patch whose cropped span [S*i + c_i, S*i + P - c_i) contains Q
// (c_0 = 0, c_i = 8), because later patches overwrite earlier ones.
__device__ __forceinline__ int stitch_owner(int Q, int cnt, int S, int P, int crop) {
  // patches i >= 1: both span ends grow with i, so only the largest i whose start is <= Q can own Q
  if (cnt > 1 && Q >= S + crop) {
    const int i = min(cnt - 1, (Q - crop) / S);
    if (Q < S * i + P - crop) return i;
  }
  // otherwise patch 0 (uncropped) if it reaches Q (matters only when step*scale < 8; never on the CLI path)
  return Q < P ? 0 : -1;
}

__global__ void patch_stitch_kernel(const float* __restrict__ patches, int cnt_h, int cnt_w, int PH,
                                    int PW, int S, int crop, int out_h, int out_w, float mul,
                                    float* __restrict__ out_f32, uint8_t* __restrict__ out_u8) {
  const size_t row_elems = (size_t)out_w * 3;
  const size_t total = (size_t)out_h * row_elems;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx % row_elems);
    const int Y = (int)(idx / row_elems);
    const int X = e / 3, c = e % 3;
    const int i = stitch_owner(Y, cnt_h, S, PH, crop);
    const int j = stitch_owner(X, cnt_w, S, PW, crop);
    float v = 0.f;
    if (i >= 0 && j >= 0) {
      const size_t n = (size_t)j * cnt_h + i;
      v = __fmul_rn(patches[((n * PH + (Y - S * i)) * PW + (X - S * j)) * 3 + c], mul);
    }
    if (out_f32) out_f32[idx] = v;
    if (out_u8) {
      // np.clip(result, 0, 255).astype('uint8'): clamp then truncate toward zero (models.py:391)
      const float cl = fminf(fmaxf(v, 0.f), 255.f);
      out_u8[idx] = (uint8_t)(int)cl;
    }
  }
}

// Sharded stitch (BASELINE config 5): this rank ran only tiles [tile_lo, tile_hi) of the column-major tile
// index (patches[0] = tile tile_lo) and writes the uint8 strip of output columns [x0, x0 + strip_w): a pixel whose
// owner tile is in the range gets its value (x mul, clip, truncate -- exactly patch_stitch_kernel's), every other
// pixel 0.  Strips of different ranks are disjoint where non-zero, so rank 0 composes them with a bitwise OR and
// the result is bit-identical to the unsharded stitch; what crosses NVLink is owned uint8 pixels (3 B) instead of
// whole fp32 patches (12 B x 2.25 overlap).  Four bytes per thread, one 4-byte store.
__global__ void __launch_bounds__(256)
patch_stitch_range_kernel(const float* __restrict__ patches, int cnt_h, int cnt_w, int PH, int PW, int S, int crop,
                          int out_h, int tile_lo, int tile_hi, int x0, int strip_w, float mul,
                          uint8_t* __restrict__ out_u8) {
  const size_t row_elems = (size_t)strip_w * 3;
  const size_t row_q = (row_elems + 3) / 4;
  const size_t total = (size_t)out_h * row_q;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int Y = (int)(idx / row_q);
    const int e0 = (int)(idx - (size_t)Y * row_q) * 4;
    const int i = stitch_owner(Y, cnt_h, S, PH, crop);
    uint8_t v[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = e0 + k;
      if (e >= (int)row_elems || i < 0) continue;
      const int X = x0 + e / 3, c = e % 3;
      const int j = stitch_owner(X, cnt_w, S, PW, crop);
      if (j < 0) continue;
      const int n = j * cnt_h + i;
      if (n < tile_lo || n >= tile_hi) continue;
      const float f = __fmul_rn(patches[(((size_t)(n - tile_lo) * PH + (Y - S * i)) * PW + (X - S * j)) * 3 + c], mul);
      v[k] = (uint8_t)(int)fminf(fmaxf(f, 0.f), 255.f);
    }
    uint8_t* dst = out_u8 + (size_t)Y * row_elems + e0;
    if (e0 + 3 < (int)row_elems && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
      *reinterpret_cast<uint32_t*>(dst) = v[0] | (v[1] << 8) | (v[2] << 16) | ((uint32_t)v[3] << 24);
    } else {
      for (int k = 0; k < 4; ++k)
        if (e0 + k < (int)row_elems) dst[k] = v[k];
    }
  }
}

// Four consecutive output elements per thread (16-byte patch load, 4-byte uint8 store).  Valid when every
// ownership boundary and row length is a multiple of 4 elements: 3*S, 3*crop, 3*PW, 3*out_w all % 4 == 0
// (true for the reference's 96/64/x4/8-px geometry); the host falls back to the scalar kernel otherwise.
__global__ void __launch_bounds__(256)
patch_stitch_vec4_kernel(const float* __restrict__ patches, int cnt_h, int cnt_w, int PH, int PW, int S, int crop,
                         int out_h, int out_w, float mul, float4* __restrict__ out_f32,
                         uint32_t* __restrict__ out_u8) {
  // one thread = 4 consecutive quads (16 output elements) of one output row: the row owner is found once, the
  // four 16-byte patch loads are issued together, the uint8 result leaves as one 16-byte store when aligned
  const int row_q = out_w * 3 / 4;
  const int row_g = (row_q + 3) / 4;
  const size_t total = (size_t)out_h * row_g;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % row_g);
    const int Y = (int)(idx / row_g);
    const int i = stitch_owner(Y, cnt_h, S, PH, crop);
    float4 v[4];
    bool have[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int q = g * 4 + k;
      have[k] = q < row_q;
      v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (have[k] && i >= 0) {
        const int e = q * 4;
        const int j = stitch_owner(e / 3, cnt_w, S, PW, crop);
        if (j >= 0) {
          const size_t n = (size_t)j * cnt_h + i;
          v[k] = *reinterpret_cast<const float4*>(patches + ((n * PH + (Y - S * i)) * PW) * 3 + (e - 3 * S * j));
        }
      }
    }
    uint32_t packed[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = make_float4(__fmul_rn(v[k].x, mul), __fmul_rn(v[k].y, mul), __fmul_rn(v[k].z, mul),
                         __fmul_rn(v[k].w, mul));
      const uint32_t b0 = (uint32_t)(int)fminf(fmaxf(v[k].x, 0.f), 255.f);
      const uint32_t b1 = (uint32_t)(int)fminf(fmaxf(v[k].y, 0.f), 255.f);
      const uint32_t b2 = (uint32_t)(int)fminf(fmaxf(v[k].z, 0.f), 255.f);
      const uint32_t b3 = (uint32_t)(int)fminf(fmaxf(v[k].w, 0.f), 255.f);
      packed[k] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    }
    const size_t q0 = (size_t)Y * row_q + (size_t)g * 4;
    if (out_f32) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (have[k]) out_f32[q0 + k] = v[k];
    }
    if (out_u8) {
      if (have[3] && (q0 & 3) == 0) {
        *reinterpret_cast<uint4*>(out_u8 + q0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (have[k]) out_u8[q0 + k] = packed[k];
      }
    }
  }
}

// ------------------------------------------------------------------ depth to space
__global__ void depth_to_space_kernel(const float* __restrict__ in, int NB, int H, int W, int C,
                                      int r, int order, float* __restrict__ out) {
  const int OH = H * r, OW = W * r;
  const size_t total = (size_t)NB * OH * OW * C;
  const int Cin = C * r * r;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    size_t q = idx / C;
    const int X = (int)(q % OW);
    q /= OW;
    const int Y = (int)(q % OH);
    const int n = (int)(q / OH);
    const int ry = Y % r, rx = X % r;
    int ch;
    if (order == 0) ch = c * r * r + rx * r + ry;
    else if (order == 1) ch = c * r * r + ry * r + rx;
    else ch = (ry * r + rx) * C + c;
    out[idx] = in[(((size_t)n * H + Y / r) * W + X / r) * Cin + ch];
  }
}

// Staged version (the one sr_depth_to_space launches): a block takes PIX consecutive input pixels of one row --
// PIX*C*r*r contiguous floats, read with full-width coalesced loads into shared memory (pixel pitch padded to an odd
// word count so the permuted reads spread over the banks) -- and writes the r output rows they map to, each a
// contiguous run of PIX*r*C floats, with 128-bit stores.  Every DRAM sector is read once and written once; the
// permutation itself happens in shared memory through a [ry][rx*C+c] -> input-channel table.
__global__ void __launch_bounds__(256)
depth_to_space_tiled_kernel(const float* __restrict__ in, int NB, int H, int W, int C, int r, int order, int PIX,
                            float* __restrict__ out) {
  extern __shared__ float d2s_smem[];
  const int Cin = C * r * r, RC = r * C, pitch = Cin | 1;
  float* tile = d2s_smem;                                        // [PIX][pitch]
  int* tbl = reinterpret_cast<int*>(d2s_smem + (size_t)PIX * pitch);  // [r][RC]
  for (int i = threadIdx.x; i < Cin; i += blockDim.x) {
    const int ry = i / RC, k = i - ry * RC, rx = k / C, c = k - rx * C;
    tbl[i] = order == 0 ? c * r * r + rx * r + ry : order == 1 ? c * r * r + ry * r + rx : (ry * r + rx) * C + c;
  }
  const int strips_per_row = (W + PIX - 1) / PIX;
  const size_t n_strips = (size_t)NB * H * strips_per_row;
  const size_t orow = (size_t)W * RC;                            // floats per output row
  for (size_t sidx = blockIdx.x; sidx < n_strips; sidx += gridDim.x) {
    const int sx = (int)(sidx % strips_per_row);
    const size_t row = sidx / strips_per_row;                    // n*H + y
    const int x0 = sx * PIX, npx = min(PIX, W - x0);
    const float* src = in + (row * W + x0) * Cin;
    const int nin = npx * Cin;
    __syncthreads();                                             // previous strip fully written out (and tbl ready)
    // (pixel, channel) of a thread's element advance by a constant per trip: one division per phase, none in the loops
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (nin & 3) == 0) {
      const int step = blockDim.x * 4, dq = step / Cin, dr = step - dq * Cin;
      int px = (threadIdx.x * 4) / Cin, ch = threadIdx.x * 4 - px * Cin;
      for (int i = threadIdx.x * 4; i < nin; i += step) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        float* t0 = tile + px * pitch + ch;
        if (ch + 3 < Cin) {
          t0[0] = v.x; t0[1] = v.y; t0[2] = v.z; t0[3] = v.w;
        } else {                                                 // the group straddles two pixels
          const float vv[4] = {v.x, v.y, v.z, v.w};
          int p2 = px, c2 = ch;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            tile[p2 * pitch + c2] = vv[e];
            if (++c2 == Cin) { c2 = 0; ++p2; }
          }
        }
        px += dq;
        ch += dr;
        if (ch >= Cin) { ch -= Cin; ++px; }
      }
    } else {
      for (int i = threadIdx.x; i < nin; i += blockDim.x) {
        const int px = i / Cin;
        tile[px * pitch + (i - px * Cin)] = src[i];
      }
    }
    __syncthreads();
    const int nout = npx * RC;                                   // floats per output row of this strip
    for (int ry = 0; ry < r; ++ry) {
      float* dst = out + (row * r + ry) * orow + (size_t)x0 * RC;
      const int* t = tbl + ry * RC;
      if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (nout & 3) == 0) {
        const int step = blockDim.x * 4, dq = step / RC, dr = step - dq * RC;
        int px = (threadIdx.x * 4) / RC, k = threadIdx.x * 4 - px * RC;
        for (int i = threadIdx.x * 4; i < nout; i += step) {
          float vv[4];
          const float* tp = tile + px * pitch;
          if (k + 3 < RC) {
            vv[0] = tp[t[k]]; vv[1] = tp[t[k + 1]]; vv[2] = tp[t[k + 2]]; vv[3] = tp[t[k + 3]];
          } else {
            int p2 = px, k2 = k;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              vv[e] = tile[p2 * pitch + t[k2]];
              if (++k2 == RC) { k2 = 0; ++p2; }
            }
          }
          *reinterpret_cast<float4*>(dst + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
          px += dq;
          k += dr;
          if (k >= RC) { k -= RC; ++px; }
        }
      } else {
        for (int i = threadIdx.x; i < nout; i += blockDim.x) {
          const int px = i / RC;
          dst[i] = tile[px * pitch + t[i - px * RC]];
        }
      }
    }
  }
}

// ------------------------------------------------------------------ weight repack
// HWIO fp32 [k*k][128][cout] -> bf16 [chunk][tap][n_pad][32]
__global__ void pack_weights_kernel(const float* __restrict__ hwio, int ntaps, int cout, int n_pad,
                                    int transpose_flip, __nv_bfloat16* __restrict__ dst) {
  const size_t total = (size_t)4 * ntaps * n_pad * 32;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx & 31);
    size_t q = idx >> 5;
    const int n = (int)(q % n_pad);
    q /= n_pad;
    const int tap = (int)(q % ntaps);
    const int chunk = (int)(q / ntaps);
    const int k_in = chunk * 32 + c;  // reduction (input-channel) index of the packed GEMM
    float v = 0.f;
    if (!transpose_flip) {
      if (n < cout) v = hwio[((size_t)tap * 128 + k_in) * cout + n];
    } else if (k_in < cout) {
      // input-gradient weights W'[t][n = ci][k = co] = W[T-1-t][ci][co]; couts beyond `cout` are zero rows of K
      v = hwio[((size_t)(ntaps - 1 - tap) * 128 + n) * cout + k_in];
    }
    dst[idx] = __float2bfloat16_rn(v);
  }
}

// tf32 variant: HWIO fp32 [k*k][128][cout] -> fp32 words rounded to tf32, [chunk of 16][tap][n_pad][16]
__device__ __forceinline__ float round_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__global__ void pack_weights_tf32_kernel(const float* __restrict__ hwio, int ntaps, int cout, int n_pad,
                                         float* __restrict__ dst) {
  const size_t total = (size_t)8 * ntaps * n_pad * 16;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx & 15);
    size_t q = idx >> 4;
    const int n = (int)(q % n_pad);
    q /= n_pad;
    const int tap = (int)(q % ntaps);
    const int chunk = (int)(q / ntaps);
    const int k_in = chunk * 16 + c;
    dst[idx] = n < cout ? round_tf32_rna(hwio[((size_t)tap * 128 + k_in) * cout + n]) : 0.f;
  }
}
__global__ void round_tf32_kernel(const float4* __restrict__ in, size_t n4, float4* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = in[i];
    out[i] = make_float4(round_tf32_rna(v.x), round_tf32_rna(v.y), round_tf32_rna(v.z), round_tf32_rna(v.w));
  }
}

// All layers in ONE launch (the optimizer step repacks 85 layers, forward and transposed: 170 tiny launches
// otherwise cost more than the all-reduce).  items[i] describes one (layer, orientation); `starts` are the prefix
// sums of the packed element counts.
__global__ void pack_weights_batched_kernel(const sr_pack_item* __restrict__ items,
                                            const unsigned long long* __restrict__ starts, int n_items) {
  const size_t total = starts[n_items];
  for (size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gidx < total;
       gidx += (size_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_items - 1;       // last item whose start <= gidx
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (starts[mid] <= gidx) lo = mid; else hi = mid - 1;
    }
    const sr_pack_item it = items[lo];
    const size_t idx = gidx - starts[lo];
    const int ntaps = it.ksize * it.ksize;
    const int n_pad = (it.cout > 16 || it.transpose_flip) ? 128 : 16;
    const int c = (int)(idx & 31);
    size_t q = idx >> 5;
    const int n = (int)(q % n_pad);
    q /= n_pad;
    const int tap = (int)(q % ntaps);
    const int chunk = (int)(q / ntaps);
    const int k_in = chunk * 32 + c;
    float v = 0.f;
    if (!it.transpose_flip) {
      if (n < it.cout) v = it.hwio[((size_t)tap * 128 + k_in) * it.cout + n];
    } else if (k_in < it.cout) {
      v = it.hwio[((size_t)(ntaps - 1 - tap) * 128 + n) * it.cout + k_in];
    }
    reinterpret_cast<__nv_bfloat16*>(it.dst)[idx] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------ direct conv (CUDA cores)
__global__ void conv_direct_kernel(const void* __restrict__ in, int in_is_bf16,
                                   const float* __restrict__ hwio, int w_round_bf16,
                                   const float* __restrict__ bias, int NB, int H, int W, int cin,
                                   int cout, int k, int same, int relu, int r, int order,
                                   float* __restrict__ out) {
  const int pad = same ? (k - 1) / 2 : 0;
  const int OH = same ? H : H - k + 1, OW = same ? W : W - k + 1;
  const size_t total = (size_t)NB * OH * OW * cout;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(idx % cout);
    size_t q = idx / cout;
    const int x = (int)(q % OW);
    q /= OW;
    const int y = (int)(q % OH);
    const int n = (int)(q / OH);
    float acc = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int iy = y + ky - pad;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = x + kx - pad;
        if (ix < 0 || ix >= W) continue;
        const size_t ib = (((size_t)n * H + iy) * W + ix) * cin;
        const float* wp = hwio + ((size_t)(ky * k + kx) * cin) * cout + co;
        for (int ci = 0; ci < cin; ++ci) {
          const float a = in_is_bf16
                              ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(in)[ib + ci])
                              : reinterpret_cast<const float*>(in)[ib + ci];
          float wv = wp[(size_t)ci * cout];
          if (w_round_bf16) wv = __bfloat162float(__float2bfloat16_rn(wv));
          acc = fmaf(a, wv, acc);
        }
      }
    }
    if (bias) acc += bias[co];
    if (relu) acc = fmaxf(acc, 0.f);
    if (r > 0) {
      const int C = cout / (r * r);
      int c, ry, rx;
      if (order == 0) { c = co / (r * r); rx = (co / r) % r; ry = co % r; }
      else if (order == 1) { c = co / (r * r); ry = (co / r) % r; rx = co % r; }
      else { c = co % C; ry = (co / C) / r; rx = (co / C) % r; }
      out[(((size_t)n * OH * r + (y * r + ry)) * OW * r + (x * r + rx)) * C + c] = acc;
    } else {
      out[idx] = acc;
    }
  }
}

// ------------------------------------------------------------------ casts / axpby / loss / adam
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, size_t n, __nv_bfloat16* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, size_t n, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}
__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float a, float b,
                             size_t n, float* __restrict__ out, __nv_bfloat16* __restrict__ out_bf16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = a * x[i];
    if (y) v = fmaf(b, y[i], v);
    if (out) out[i] = v;
    if (out_bf16) out_bf16[i] = __float2bfloat16_rn(v);
  }
}

__global__ void mse_loss_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                     size_t n, float inv_total2, float* __restrict__ grad,
                                     double* __restrict__ loss_sum) {
  double local = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = pred[i] - target[i];
    local += (double)d * (double)d;
    if (grad) grad[i] = d * inv_total2;
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double sm[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = local;
  __syncthreads();
  if (wid == 0) {
    double v = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && loss_sum) atomicAdd(loss_sum, v);
  }
}

// Keras 2 Adam (keras/optimizers.py Adam.get_updates): lr_t = lr*sqrt(1-b2^t)/(1-b1^t)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, float lr_t, float b1, float b2, float eps,
                            float gscale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace
}  // namespace sr

using namespace sr;

extern "C" int sr_head1x1_fwd(const float* in, const float* w, const float* bias, size_t npix,
                              void* out_bf16, float* out_f32, void* stream) {
  if (!in || !w || (!out_bf16 && !out_f32)) return set_error(SR_ERR_INVALID, "sr_head1x1_fwd: null pointer");
  if (npix == 0) return SR_OK;
  head1x1_kernel<<<grid_for(npix * 16, kBlock * 4, 148 * 8), kBlock, 0, as_stream(stream)>>>(
      in, w, bias, npix, reinterpret_cast<uint4*>(out_bf16), reinterpret_cast<float4*>(out_f32));
  return check_launch("head1x1_kernel");
}

static int bilinear4_launch(const void* in, int in_is_bf16, const int* src_index, int NB, int H, int W, int C,
                            int CH, int CW, void* out_bf16, float* out_f32, void* stream) {
  const size_t total = (size_t)NB * CH * CW * (C / 8);
  const unsigned g = grid_for(total, kBlock, 148 * 32);
  if (in_is_bf16)
    bilinear4_fwd_kernel<true><<<g, kBlock, 0, as_stream(stream)>>>(
        in, src_index, NB, H, W, C, CH, CW, reinterpret_cast<uint4*>(out_bf16), reinterpret_cast<float4*>(out_f32));
  else
    bilinear4_fwd_kernel<false><<<g, kBlock, 0, as_stream(stream)>>>(
        in, src_index, NB, H, W, C, CH, CW, reinterpret_cast<uint4*>(out_bf16), reinterpret_cast<float4*>(out_f32));
  return check_launch("bilinear4_fwd_kernel");
}

extern "C" int sr_bilinear4_fwd(const void* in, int in_is_bf16, int NB, int H, int W, int C,
                                void* out_bf16, float* out_f32, void* stream) {
  if (!in || (!out_bf16 && !out_f32)) return set_error(SR_ERR_INVALID, "sr_bilinear4_fwd: null pointer");
  if (C % 8 != 0) return set_error(SR_ERR_UNSUPPORTED, "sr_bilinear4_fwd: C must be a multiple of 8");
  if (NB < 1 || H < 1 || W < 1) return set_error(SR_ERR_INVALID, "sr_bilinear4_fwd: empty tensor");
  return bilinear4_launch(in, in_is_bf16, nullptr, NB, H, W, C, H, W, out_bf16, out_f32, stream);
}

// x2 with the same legacy sampling (src = dst * 0.5): Lambda(resize2bil) of Difvdsr4 (models.py:932-940, 1046, 1053)
extern "C" int sr_bilinear2_fwd(const void* in, int in_is_bf16, int NB, int H, int W, int C, void* out_bf16,
                                float* out_f32, void* stream) {
  if (!in || (!out_bf16 && !out_f32)) return set_error(SR_ERR_INVALID, "sr_bilinear2_fwd: null pointer");
  if (C % 8 != 0) return set_error(SR_ERR_UNSUPPORTED, "sr_bilinear2_fwd: C must be a multiple of 8");
  if (NB < 1 || H < 1 || W < 1) return set_error(SR_ERR_INVALID, "sr_bilinear2_fwd: empty tensor");
  const size_t total = (size_t)NB * H * W * (C / 8);
  const unsigned g = grid_for(total, kBlock, 148 * 32);
  if (in_is_bf16)
    bilinear4_fwd_kernel<true, 2><<<g, kBlock, 0, as_stream(stream)>>>(
        in, nullptr, NB, H, W, C, H, W, reinterpret_cast<uint4*>(out_bf16), reinterpret_cast<float4*>(out_f32));
  else
    bilinear4_fwd_kernel<false, 2><<<g, kBlock, 0, as_stream(stream)>>>(
        in, nullptr, NB, H, W, C, H, W, reinterpret_cast<uint4*>(out_bf16), reinterpret_cast<float4*>(out_f32));
  return check_launch("bilinear2_fwd_kernel");
}

extern "C" int sr_bilinear4_crop_fwd(const void* in, int in_is_bf16, const int* src_index, int n_out, int H,
                                     int W, int C, int out_h, int out_w, void* out_bf16, float* out_f32,
                                     void* stream) {
  if (!in || (!out_bf16 && !out_f32)) return set_error(SR_ERR_INVALID, "sr_bilinear4_crop_fwd: null pointer");
  if (C % 8 != 0) return set_error(SR_ERR_UNSUPPORTED, "sr_bilinear4_crop_fwd: C must be a multiple of 8");
  if (n_out < 1 || H < 1 || W < 1) return set_error(SR_ERR_INVALID, "sr_bilinear4_crop_fwd: empty tensor");
  if (out_h < 4 || out_w < 4 || out_h % 4 || out_w % 4 || out_h > 4 * H || out_w > 4 * W)
    return set_error(SR_ERR_INVALID, "sr_bilinear4_crop_fwd: output extent must be a multiple of 4 within 4H x 4W");
  return bilinear4_launch(in, in_is_bf16, src_index, n_out, H, W, C, out_h / 4, out_w / 4, out_bf16, out_f32,
                          stream);
}

extern "C" int sr_bilinear4_bwd(const float* gout, int NB, int H, int W, int C, float* gin, void* stream) {
  if (!gout || !gin) return set_error(SR_ERR_INVALID, "sr_bilinear4_bwd: null pointer");
  if (C % 4 != 0) return set_error(SR_ERR_UNSUPPORTED, "sr_bilinear4_bwd: C must be a multiple of 4");
  const size_t total = (size_t)NB * H * W * (C / 4);
  bilinear4_bwd_kernel<4><<<grid_for(total, kBlock, 148 * 32), kBlock, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(gout), NB, H, W, C, reinterpret_cast<float4*>(gin));
  return check_launch("bilinear4_bwd_kernel");
}

extern "C" int sr_bilinear2_bwd(const float* gout, int NB, int H, int W, int C, float* gin, void* stream) {
  if (!gout || !gin) return set_error(SR_ERR_INVALID, "sr_bilinear2_bwd: null pointer");
  if (C % 4 != 0) return set_error(SR_ERR_UNSUPPORTED, "sr_bilinear2_bwd: C must be a multiple of 4");
  const size_t total = (size_t)NB * H * W * (C / 4);
  bilinear4_bwd_kernel<2><<<grid_for(total, kBlock, 148 * 32), kBlock, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(gout), NB, H, W, C, reinterpret_cast<float4*>(gin));
  return check_launch("bilinear2_bwd_kernel");
}

extern "C" int sr_patch_count(int dim, int patch, int step) {
  if (step <= 0 || patch <= 0) return 0;
  const int lim = dim - patch;  // positions x with 0 <= x < lim and x % step == 0
  if (lim <= 0) return 0;
  return (lim - 1) / step + 1;
}

extern "C" int sr_canvas_size(int h, int w, int patch, int step, int* canvas_h, int* canvas_w) {
  if (!canvas_h || !canvas_w || h < 1 || w < 1 || patch < 1 || step < 1)
    return set_error(SR_ERR_INVALID, "sr_canvas_size: bad argument");
  int ch = h + patch, cw = w + patch;
  if (cw % step != 0 || ch % step != 0) {
    // models.py:250-252: int((x/step)+1)*step on BOTH dims
    cw = (cw / step + 1) * step;
    ch = (ch / step + 1) * step;
  }
  *canvas_h = ch;
  *canvas_w = cw;
  return SR_OK;
}

extern "C" int sr_patch_gather_u8_batched(const uint8_t* imgs, int n_img, size_t img_stride, int h, int w,
                                          int canvas_h, int canvas_w, int ph, int pw, int step, float divisor,
                                          float* out_f32, void* stream) {
  if (!imgs || !out_f32) return set_error(SR_ERR_INVALID, "sr_patch_gather_u8: null pointer");
  if (n_img < 1 || h < 1 || w < 1 || step < 1 || divisor == 0.f) return set_error(SR_ERR_INVALID, "sr_patch_gather_u8: bad size");
  if (ph > canvas_h) return set_error(SR_ERR_INVALID, "Height of the patch should be less than the height of the image.");
  if (pw > canvas_w) return set_error(SR_ERR_INVALID, "Width of the patch should be less than the width of the image.");
  const int cnt_h = sr_patch_count(canvas_h, ph, step), cnt_w = sr_patch_count(canvas_w, pw, step);
  const size_t per_img = (size_t)cnt_h * cnt_w * ph * pw * 3;
  if (per_img == 0) return SR_OK;
  const size_t n_patches = (size_t)n_img * cnt_h * cnt_w;
  if ((pw * 3) % 4 == 0 && (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0 && n_patches < 0x7fffffffu) {
    const unsigned grid = grid_for(n_patches, 1, 148 * 8);
    float4* o4 = reinterpret_cast<float4*>(out_f32);
    if (divisor == 255.f)
      patch_gather_u8_rows_kernel<1><<<grid, kBlock, 0, as_stream(stream)>>>(imgs, n_img, img_stride, h, w, cnt_h, cnt_w,
                                                                            ph, pw, step, divisor, o4);
    else if (divisor == 1.f)
      patch_gather_u8_rows_kernel<2><<<grid, kBlock, 0, as_stream(stream)>>>(imgs, n_img, img_stride, h, w, cnt_h, cnt_w,
                                                                            ph, pw, step, divisor, o4);
    else
      patch_gather_u8_rows_kernel<0><<<grid, kBlock, 0, as_stream(stream)>>>(imgs, n_img, img_stride, h, w, cnt_h, cnt_w,
                                                                            ph, pw, step, divisor, o4);
    return check_launch("patch_gather_u8_rows_kernel");
  }
  for (int m = 0; m < n_img; ++m) {     // odd patch widths: the scalar kernel, one image per launch
    patch_gather_kernel<true><<<grid_for(per_img, kBlock, 148 * 32), kBlock, 0, as_stream(stream)>>>(
        imgs + (size_t)m * img_stride, h, w, canvas_w, cnt_h, cnt_w, ph, pw, step, divisor, out_f32 + (size_t)m * per_img);
    const int rc = check_launch("patch_gather_kernel<u8>");
    if (rc != SR_OK) return rc;
  }
  return SR_OK;
}

extern "C" int sr_patch_gather_u8(const uint8_t* img, int h, int w, int canvas_h, int canvas_w, int ph,
                                  int pw, int step, float divisor, float* out_f32, void* stream) {
  return sr_patch_gather_u8_batched(img, 1, 0, h, w, canvas_h, canvas_w, ph, pw, step, divisor, out_f32, stream);
}

extern "C" int sr_batch_gather_u8(const uint8_t* data, size_t item_bytes, size_t n_items, const long long* index,
                                  int n, float divisor, float* out_f32, void* stream) {
  if (!data || !index || !out_f32) return set_error(SR_ERR_INVALID, "sr_batch_gather_u8: null pointer");
  if (n < 0 || item_bytes == 0 || divisor == 0.f) return set_error(SR_ERR_INVALID, "sr_batch_gather_u8: bad size");
  if (n == 0) return SR_OK;
  (void)n_items;  // the index values are device data: the caller guarantees 0 <= index[i] < n_items
  if ((item_bytes & 3) != 0)
    return set_error(SR_ERR_UNSUPPORTED, "sr_batch_gather_u8: item_bytes must be a multiple of 4");
  if ((reinterpret_cast<uintptr_t>(data) & 3) != 0 || (reinterpret_cast<uintptr_t>(out_f32) & 15) != 0)
    return set_error(SR_ERR_INVALID, "sr_batch_gather_u8: data must be 4-byte and out 16-byte aligned");
  const unsigned grid = grid_for((size_t)n * (item_bytes >> 2), kBlock * 4, 148 * 8);
  if (divisor == 255.f)
    batch_gather_u8_kernel<true><<<grid, kBlock, 0, as_stream(stream)>>>(data, item_bytes, index, n, divisor, out_f32);
  else
    batch_gather_u8_kernel<false><<<grid, kBlock, 0, as_stream(stream)>>>(data, item_bytes, index, n, divisor, out_f32);
  return check_launch("batch_gather_u8_kernel");
}

extern "C" int sr_patch_gather_f32(const float* canvas, int canvas_h, int canvas_w, int ph, int pw,
                                   int step, float* out_f32, void* stream) {
  if (!canvas || !out_f32) return set_error(SR_ERR_INVALID, "sr_patch_gather_f32: null pointer");
  if (ph > canvas_h) return set_error(SR_ERR_INVALID, "Height of the patch should be less than the height of the image.");
  if (pw > canvas_w) return set_error(SR_ERR_INVALID, "Width of the patch should be less than the width of the image.");
  const int cnt_h = sr_patch_count(canvas_h, ph, step), cnt_w = sr_patch_count(canvas_w, pw, step);
  const size_t total = (size_t)cnt_h * cnt_w * ph * pw * 3;
  if (total == 0) return SR_OK;
  patch_gather_kernel<false><<<grid_for(total, kBlock, 148 * 32), kBlock, 0, as_stream(stream)>>>(
      canvas, canvas_h, canvas_w, canvas_w, cnt_h, cnt_w, ph, pw, step, 1.f, out_f32);
  return check_launch("patch_gather_kernel<f32>");
}

extern "C" int sr_patch_stitch(const float* patches, int cnt_h, int cnt_w, int ph, int pw, int step,
                               int scale, int canvas_h, int canvas_w, float mul, float* out_f32,
                               uint8_t* out_u8, void* stream) {
  if (!patches || (!out_f32 && !out_u8)) return set_error(SR_ERR_INVALID, "sr_patch_stitch: null pointer");
  if (cnt_h < 1 || cnt_w < 1 || scale < 1) return set_error(SR_ERR_INVALID, "sr_patch_stitch: bad counts");
  const int out_h = canvas_h * scale, out_w = canvas_w * scale;
  const size_t total = (size_t)out_h * out_w * 3;
  const int S = step * scale, PW = pw * scale, crop = 8;
  const bool vec_ok = (3 * S) % 4 == 0 && (3 * crop) % 4 == 0 && (3 * PW) % 4 == 0 && (3 * out_w) % 4 == 0 &&
                      ((reinterpret_cast<uintptr_t>(patches) | reinterpret_cast<uintptr_t>(out_f32)) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(out_u8) & 15) == 0;
  if (vec_ok) {
    patch_stitch_vec4_kernel<<<grid_for((size_t)out_h * ((out_w * 3 / 4 + 3) / 4), kBlock, 148 * 32), kBlock, 0, as_stream(stream)>>>(
        patches, cnt_h, cnt_w, ph * scale, PW, S, crop, out_h, out_w, mul, reinterpret_cast<float4*>(out_f32),
        reinterpret_cast<uint32_t*>(out_u8));
    return check_launch("patch_stitch_vec4_kernel");
  }
  patch_stitch_kernel<<<grid_for(total, kBlock, 148 * 32), kBlock, 0, as_stream(stream)>>>(
      patches, cnt_h, cnt_w, ph * scale, PW, S, crop, out_h, out_w, mul, out_f32, out_u8);
  return check_launch("patch_stitch_kernel");
}

extern "C" int sr_patch_stitch_range(const float* patches, int cnt_h, int cnt_w, int ph, int pw, int step,
                                     int scale, int canvas_h, int tile_lo, int tile_hi, int x0, int strip_w,
                                     float mul, uint8_t* out_u8, void* stream) {
  if (!patches || !out_u8) return set_error(SR_ERR_INVALID, "sr_patch_stitch_range: null pointer");
  if (cnt_h < 1 || cnt_w < 1 || scale < 1 || tile_lo < 0 || tile_hi < tile_lo || tile_hi > cnt_h * cnt_w ||
      x0 < 0 || strip_w < 1)
    return set_error(SR_ERR_INVALID, "sr_patch_stitch_range: bad counts / tile range / strip");
  const int out_h = canvas_h * scale;
  const size_t total = (size_t)out_h * (((size_t)strip_w * 3 + 3) / 4);
  patch_stitch_range_kernel<<<grid_for(total, kBlock, 148 * 32), kBlock, 0, as_stream(stream)>>>(
      patches, cnt_h, cnt_w, ph * scale, pw * scale, step * scale, 8, out_h, tile_lo, tile_hi, x0, strip_w, mul, out_u8);
  return check_launch("patch_stitch_range_kernel");
}

extern "C" int sr_depth_to_space(const float* in, int NB, int H, int W, int C, int r, int order,
                                 float* out, void* stream) {
  if (!in || !out) return set_error(SR_ERR_INVALID, "sr_depth_to_space: null pointer");
  if (r < 1 || C < 1 || order < 0 || order > 2) return set_error(SR_ERR_INVALID, "sr_depth_to_space: bad r/C/order");
  const size_t total = (size_t)NB * H * r * W * r * C;
  if (total == 0) return SR_OK;
  const int Cin = C * r * r, pitch = Cin | 1;
  if ((size_t)pitch * 4 + (size_t)Cin * 4 <= 40 * 1024) {
    // strips of <= ~24 KB of input: several blocks per SM keep loads and stores of different strips in flight
    int PIX = (int)std::min<size_t>((size_t)W, std::max<size_t>(1, (24 * 1024) / ((size_t)pitch * 4)));
    if (PIX >= 4) PIX &= ~3;
    const size_t smem = (size_t)PIX * pitch * 4 + (size_t)Cin * 4;
    const size_t n_strips = (size_t)NB * H * ((W + PIX - 1) / PIX);
    depth_to_space_tiled_kernel<<<(unsigned)std::min<size_t>(n_strips, 148 * 8), kBlock, smem, as_stream(stream)>>>(
        in, NB, H, W, C, r, order, PIX, out);
    return check_launch("depth_to_space_tiled_kernel");
  }
  depth_to_space_kernel<<<grid_for(total, kBlock, 148 * 32), kBlock, 0, as_stream(stream)>>>(
      in, NB, H, W, C, r, order, out);
  return check_launch("depth_to_space_kernel");
}

extern "C" size_t sr_packed_weight_bytes(int ksize, int cout) {
  const int n_pad = cout > 16 ? 128 : 16;
  return (size_t)4 * ksize * ksize * n_pad * 32 * 2;
}

extern "C" int sr_pack_conv_weights(const float* hwio, int ksize, int cout, int transpose_flip,
                                    void* dst, void* stream) {
  if (!hwio || !dst) return set_error(SR_ERR_INVALID, "sr_pack_conv_weights: null pointer");
  if (cout < 1 || cout > 128) return set_error(SR_ERR_UNSUPPORTED, "sr_pack_conv_weights: cout must be 1..128");
  const int n_pad = (cout > 16 || transpose_flip) ? 128 : 16;  // the input-gradient conv always has 128 outputs
  const size_t total = (size_t)4 * ksize * ksize * n_pad * 32;
  pack_weights_kernel<<<grid_for(total, kBlock), kBlock, 0, as_stream(stream)>>>(
      hwio, ksize * ksize, cout, n_pad, transpose_flip, reinterpret_cast<__nv_bfloat16*>(dst));
  return check_launch("pack_weights_kernel");
}

extern "C" size_t sr_packed_weight_bytes_tf32(int ksize, int cout) {
  const int n_pad = cout > 16 ? 128 : 16;
  return (size_t)8 * ksize * ksize * n_pad * 16 * 4;
}

extern "C" int sr_pack_conv_weights_tf32(const float* hwio, int ksize, int cout, void* dst, void* stream) {
  if (!hwio || !dst) return set_error(SR_ERR_INVALID, "sr_pack_conv_weights_tf32: null pointer");
  if (cout < 1 || cout > 128) return set_error(SR_ERR_UNSUPPORTED, "sr_pack_conv_weights_tf32: cout must be 1..128");
  const int n_pad = cout > 16 ? 128 : 16;
  const size_t total = (size_t)8 * ksize * ksize * n_pad * 16;
  pack_weights_tf32_kernel<<<grid_for(total, kBlock), kBlock, 0, as_stream(stream)>>>(
      hwio, ksize * ksize, cout, n_pad, reinterpret_cast<float*>(dst));
  return check_launch("pack_weights_tf32_kernel");
}

extern "C" int sr_round_tf32(const float* in, size_t n, float* out, void* stream) {
  if (n == 0) return SR_OK;
  if (!in || !out) return set_error(SR_ERR_INVALID, "sr_round_tf32: null pointer");
  if (n % 4 != 0 || (reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 != 0)
    return set_error(SR_ERR_UNSUPPORTED, "sr_round_tf32: n % 4 == 0 and 16-byte aligned pointers");
  round_tf32_kernel<<<grid_for(n / 4, kBlock, 148 * 16), kBlock, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(in), n / 4, reinterpret_cast<float4*>(out));
  return check_launch("round_tf32_kernel");
}

extern "C" int sr_pack_conv_weights_batched(const sr_pack_item* items_dev, const unsigned long long* starts_dev,
                                            int n_items, size_t total_elems, void* stream) {
  if (!items_dev || !starts_dev) return set_error(SR_ERR_INVALID, "sr_pack_conv_weights_batched: null pointer");
  if (n_items < 1 || total_elems == 0) return SR_OK;
  pack_weights_batched_kernel<<<grid_for(total_elems, kBlock, 148 * 16), kBlock, 0, as_stream(stream)>>>(
      items_dev, starts_dev, n_items);
  return check_launch("pack_weights_batched_kernel");
}

extern "C" int sr_conv2d_direct(const void* in, int in_is_bf16, const float* hwio, int w_round_bf16,
                                const float* bias, int NB, int H, int W, int cin, int cout, int ksize,
                                int same_padding, int relu, int shuffle_r, int shuffle_order, float* out,
                                void* stream) {
  if (!in || !hwio || !out) return set_error(SR_ERR_INVALID, "sr_conv2d_direct: null pointer");
  if (shuffle_r > 0 && cout % (shuffle_r * shuffle_r) != 0)
    return set_error(SR_ERR_INVALID, "sr_conv2d_direct: cout not divisible by r*r");
  const int OH = same_padding ? H : H - ksize + 1, OW = same_padding ? W : W - ksize + 1;
  if (OH < 1 || OW < 1) return set_error(SR_ERR_INVALID, "sr_conv2d_direct: kernel larger than image");
  const size_t total = (size_t)NB * OH * OW * cout;
  conv_direct_kernel<<<grid_for(total, kBlock, 148 * 32), kBlock, 0, as_stream(stream)>>>(
      in, in_is_bf16, hwio, w_round_bf16, bias, NB, H, W, cin, cout, ksize, same_padding, relu,
      shuffle_r, shuffle_order, out);
  return check_launch("conv_direct_kernel");
}

extern "C" int sr_cast_f32_to_bf16(const float* in, size_t n, void* out_bf16, void* stream) {
  if (!in || !out_bf16) return set_error(SR_ERR_INVALID, "sr_cast_f32_to_bf16: null pointer");
  if (n == 0) return SR_OK;
  cast_f32_bf16_kernel<<<grid_for(n, kBlock), kBlock, 0, as_stream(stream)>>>(in, n, reinterpret_cast<__nv_bfloat16*>(out_bf16));
  return check_launch("cast_f32_bf16_kernel");
}
extern "C" int sr_cast_bf16_to_f32(const void* in_bf16, size_t n, float* out, void* stream) {
  if (!in_bf16 || !out) return set_error(SR_ERR_INVALID, "sr_cast_bf16_to_f32: null pointer");
  if (n == 0) return SR_OK;
  cast_bf16_f32_kernel<<<grid_for(n, kBlock), kBlock, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(in_bf16), n, out);
  return check_launch("cast_bf16_f32_kernel");
}
extern "C" int sr_axpby_f32(const float* x, const float* y, float a, float b, size_t n, float* out,
                            void* out_bf16, void* stream) {
  if (!x || (!out && !out_bf16)) return set_error(SR_ERR_INVALID, "sr_axpby_f32: null pointer");
  if (n == 0) return SR_OK;
  axpby_kernel<<<grid_for(n, kBlock), kBlock, 0, as_stream(stream)>>>(x, y, a, b, n, out, reinterpret_cast<__nv_bfloat16*>(out_bf16));
  return check_launch("axpby_kernel");
}

extern "C" int sr_mse_loss_grad(const float* pred, const float* target, size_t n, size_t n_total,
                                float* grad, double* loss_sum, void* stream) {
  if (!pred || !target) return set_error(SR_ERR_INVALID, "sr_mse_loss_grad: null pointer");
  if (n == 0) return SR_OK;
  const float inv2 = (float)(2.0 / (double)n_total);
  mse_loss_grad_kernel<<<grid_for(n, kBlock, 148 * 8), kBlock, 0, as_stream(stream)>>>(pred, target, n, inv2, grad, loss_sum);
  return check_launch("mse_loss_grad_kernel");
}

extern "C" int sr_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                            float beta2, float eps, int t, float grad_scale, void* stream) {
  if (!p || !g || !m || !v) return set_error(SR_ERR_INVALID, "sr_adam_step: null pointer");
  if (t < 1) return set_error(SR_ERR_INVALID, "sr_adam_step: t must be >= 1");
  if (n == 0) return SR_OK;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t));
  adam_kernel<<<grid_for(n, kBlock), kBlock, 0, as_stream(stream)>>>(p, g, m, v, n, (float)lr_t, beta1, beta2, eps, grad_scale);
  return check_launch("adam_kernel");
}
