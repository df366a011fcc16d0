// Bandwidth-bound pieces of the training step (models.py:131-157 fit; compile(loss='mse', Adam) models.py:1212-1213)
// that sit around the tensor-core dgrad / wgrad kernels: loss gradient at the tail, bias gradients (column
// sums), first-layer (1x1, 3->128) weight gradient.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.h"

namespace sr {
namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// loss_sum += sum (pred-target)^2 ; g[pix][0..C) = 2*(pred-target)/n_total where pred > 0 (the ReLU of the last
// Conv2D, models.py:1199), written as bf16 rows of 128 channels (channels >= C are zero) = the operand layout of
// the dgrad / wgrad kernels.  16 threads per pixel, one 16-byte store each.
__global__ void mse_tail_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                     size_t npix, int C, float inv_total2, uint4* __restrict__ g128,
                                     double* __restrict__ loss_sum) {
  double local = 0.0;
  const size_t total = npix * 16;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i >> 4;
    const int q = (int)(i & 15);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (q * 8 < C) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = q * 8 + e;
        v[e] = 0.f;
        if (c < C) {
          const float p = pred[pix * C + c];
          const float d = p - target[pix * C + c];
          local += (double)d * (double)d;
          v[e] = p > 0.f ? d * inv_total2 : 0.f;
        }
      }
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
      o = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                     *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
    }
    g128[i] = o;
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

// Same loss gradient, laid out as the im2col of the 3x3 tail conv's backward: channel j = (ky*3+kx)*3 + co of pixel
// (y,x) holds g3[y-ky+1][x-kx+1][co] (zero outside the image), channels 27..127 zero.  With this tensor both tail
// gradients are 1x1 problems for the 128-wide tensor-core kernels:
//   dgrad: gx[pix][ci]  = sum_j A[pix][j] * W[ky][kx][ci][co]                (1x1 conv, K = 27 padded to 128)
//   wgrad: dW[ky][kx][ci][co] = sum_pix x[pix][ci] * A[pix][j]               (k = 1 filter gradient)
// loss_sum += sum (pred-target)^2 and db3[co] += sum g3[..][co] (each pixel counted once: by the block that owns it).
// One block per strip of kColStrip pixels x kColRows image rows: the kColRows + 2 gradient rows around it are staged
// in shared memory as bf16-rounded floats (with a 1-pixel halo: 1.25x redundant reads instead of the 3x of one-row
// blocks, and the two barriers / the block reduction / the four atomics are paid once per 256 KB written instead of
// once per 32 KB), then 16 threads per pixel write its 256-byte row.
constexpr int kColStrip = 128;
constexpr int kColRows = 8;

__global__ void __launch_bounds__(kBlock)
mse_tail_grad_col_kernel(const float* __restrict__ pred, const float* __restrict__ target, int H, int W, int strips,
                         int row_blocks, float inv_total2, uint4* __restrict__ a128, double* __restrict__ loss_sum,
                         float* __restrict__ db3) {
  constexpr int kRow = (kColStrip + 2) * 3;
  __shared__ float g[(kColRows + 2) * kRow];
  __shared__ double red_loss[kBlock / 32];
  __shared__ float red_db[kBlock / 32][3];
  const int strip = blockIdx.x % strips;
  const int yb = (blockIdx.x / strips) % row_blocks;
  const size_t n = blockIdx.x / ((size_t)strips * row_blocks);
  const int x0 = strip * kColStrip, y0 = yb * kColRows;
  const int npx = min(kColStrip, W - x0), nrows = min(kColRows, H - y0);
  double loss = 0.0;
  float db[3] = {0.f, 0.f, 0.f};
  for (int r = 0; r < nrows + 2; ++r) {
    const int sy = y0 - 1 + r;
    const bool row_in = sy >= 0 && sy < H;
    const bool own_row = r >= 1 && r <= nrows;
    const float* prow = pred + (((size_t)n * H + (row_in ? sy : 0)) * W) * 3;
    const float* trow = target + (((size_t)n * H + (row_in ? sy : 0)) * W) * 3;
    for (int i = threadIdx.x; i < kRow; i += kBlock) {
      const int xx = i / 3, c = i - xx * 3;
      const int sx = x0 - 1 + xx;
      float v = 0.f;
      if (row_in && sx >= 0 && sx < W && xx <= npx + 1) {
        const float p = prow[(size_t)sx * 3 + c];
        const float d = p - trow[(size_t)sx * 3 + c];
        v = __bfloat162float(__float2bfloat16_rn(p > 0.f ? d * inv_total2 : 0.f));
        if (own_row && xx >= 1 && xx <= npx) {
          loss += (double)d * (double)d;
          db[0] += c == 0 ? v : 0.f;
          db[1] += c == 1 ? v : 0.f;
          db[2] += c == 2 ? v : 0.f;
        }
      }
      g[r * kRow + i] = v;
    }
  }
  // element offsets of this thread's 8 channels (fixed: kBlock is a multiple of 16)
  const int q = threadIdx.x & 15;
  int off[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int j = q * 8 + e;
    const int tap = j / 3, co = j - tap * 3;
    const int ky = tap / 3, kx = tap - ky * 3;
    off[e] = j < 27 ? ((2 - ky) * (kColStrip + 2) + (2 - kx)) * 3 + co : -1;
  }
  __syncthreads();
  for (int rr = 0; rr < nrows; ++rr) {
    const size_t base = (((size_t)n * H + y0 + rr) * W + x0) * 16;
    const float* gr = g + rr * kRow;
    for (int i = threadIdx.x; i < npx * 16; i += kBlock) {
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (q < 4) {
        const int px3 = (i >> 4) * 3;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = off[e] >= 0 ? gr[off[e] + px3] : 0.f;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
        o = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                       *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
      }
      a128[base + i] = o;
    }
  }
  // block reductions: loss (fp64) and the three bias-gradient sums
  for (int o = 16; o > 0; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o);
#pragma unroll
    for (int c = 0; c < 3; ++c) db[c] += __shfl_xor_sync(0xffffffffu, db[c], o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    red_loss[wid] = loss;
    red_db[wid][0] = db[0]; red_db[wid][1] = db[1]; red_db[wid][2] = db[2];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    for (int w = 0; w < kBlock / 32; ++w) {
      l += red_loss[w]; b0 += red_db[w][0]; b1 += red_db[w][1]; b2 += red_db[w][2];
    }
    if (loss_sum) atomicAdd(loss_sum, l);
    if (db3) { atomicAdd(db3, b0); atomicAdd(db3 + 1, b1); atomicAdd(db3 + 2, b2); }
  }
}

// out[c] += scale * sum_pix g[pix][c], g bf16 [npix][128].  16 threads per pixel row (8 channels each), 16 rows
// per block pass; block-level shared-memory reduction, one fp32 atomic per channel per block.
__global__ void colsum_bf16_kernel(const uint4* __restrict__ g, size_t npix, float scale, float* __restrict__ out) {
  const int q = threadIdx.x & 15, r = threadIdx.x >> 4;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (size_t pix = (size_t)blockIdx.x * 16 + r; pix < npix; pix += (size_t)gridDim.x * 16) {
    const uint4 v = g[pix * 16 + q];
    acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x);
    acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
    acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z);
    acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
  }
  __shared__ float sm[16][129];
#pragma unroll
  for (int e = 0; e < 8; ++e) sm[r][q * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += sm[i][threadIdx.x];
    atomicAdd(out + threadIdx.x, scale * s);
  }
}

// First layer backward (Convolution2D(128,(1,1),relu,name='level1'), models.py:1177):
//   g0 = g * (act > 0);  dW[c3][co] += sum_pix x[pix][c3] * g0[pix][co];  db[co] += sum_pix g0[pix][co].
// One thread per output channel, blocks stride over pixels.
__global__ void head1x1_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ act,
                                   const float* __restrict__ g_f32, const __nv_bfloat16* __restrict__ g_bf16,
                                   size_t npix, float* __restrict__ dw, float* __restrict__ db) {
  const int co = threadIdx.x;  // 128 threads
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, ab = 0.f;
  for (size_t pix = blockIdx.x; pix < npix; pix += gridDim.x) {
    const size_t o = pix * 128 + co;
    float gv = g_f32 ? g_f32[o] : __bfloat162float(g_bf16[o]);
    if (!(__bfloat162float(act[o]) > 0.f)) gv = 0.f;
    const float x0 = x[pix * 3], x1 = x[pix * 3 + 1], x2 = x[pix * 3 + 2];
    a0 = fmaf(x0, gv, a0);
    a1 = fmaf(x1, gv, a1);
    a2 = fmaf(x2, gv, a2);
    ab += gv;
  }
  atomicAdd(dw + co, a0);
  atomicAdd(dw + 128 + co, a1);
  atomicAdd(dw + 256 + co, a2);
  atomicAdd(db + co, ab);
}

}  // namespace
}  // namespace sr

using namespace sr;

extern "C" int sr_mse_tail_grad(const float* pred, const float* target, size_t npix, int channels,
                                size_t n_total, void* g128_bf16, double* loss_sum, void* stream) {
  if (!pred || !target || !g128_bf16) return set_error(SR_ERR_INVALID, "sr_mse_tail_grad: null pointer");
  if (channels < 1 || channels > 128) return set_error(SR_ERR_INVALID, "sr_mse_tail_grad: channels must be 1..128");
  if (npix == 0) return SR_OK;
  const float inv2 = (float)(2.0 / (double)n_total);
  mse_tail_grad_kernel<<<grid_for(npix * 16, kBlock, 148 * 16), kBlock, 0, as_stream(stream)>>>(
      pred, target, npix, channels, inv2, reinterpret_cast<uint4*>(g128_bf16), loss_sum);
  return check_launch("mse_tail_grad_kernel");
}

extern "C" int sr_mse_tail_grad_col(const float* pred, const float* target, int NB, int H, int W, size_t n_total,
                                    void* a128_bf16, double* loss_sum, float* db3, void* stream) {
  if (!pred || !target || !a128_bf16) return set_error(SR_ERR_INVALID, "sr_mse_tail_grad_col: null pointer");
  if (NB < 1 || H < 1 || W < 1) return set_error(SR_ERR_INVALID, "sr_mse_tail_grad_col: empty tensor");
  const float inv2 = (float)(2.0 / (double)n_total);
  const int strips = (W + kColStrip - 1) / kColStrip;
  const int row_blocks = (H + kColRows - 1) / kColRows;
  const size_t blocks = (size_t)NB * row_blocks * strips;
  if (blocks > 0x7fffffffull) return set_error(SR_ERR_INVALID, "sr_mse_tail_grad_col: tensor too large");
  mse_tail_grad_col_kernel<<<(unsigned)blocks, kBlock, 0, as_stream(stream)>>>(
      pred, target, H, W, strips, row_blocks, inv2, reinterpret_cast<uint4*>(a128_bf16), loss_sum, db3);
  return check_launch("mse_tail_grad_col_kernel");
}

extern "C" int sr_colsum_bf16(const void* g_bf16, size_t npix, float scale, float* out, void* stream) {
  if (!g_bf16 || !out) return set_error(SR_ERR_INVALID, "sr_colsum_bf16: null pointer");
  if (npix == 0) return SR_OK;
  colsum_bf16_kernel<<<grid_for(npix, 16, 148 * 8), kBlock, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(g_bf16), npix, scale, out);
  return check_launch("colsum_bf16_kernel");
}

extern "C" int sr_head1x1_bwd(const float* x, const void* act_bf16, const float* g_f32, const void* g_bf16,
                              size_t npix, float* dw, float* db, void* stream) {
  if (!x || !act_bf16 || (!g_f32 && !g_bf16) || !dw || !db)
    return set_error(SR_ERR_INVALID, "sr_head1x1_bwd: null pointer");
  if (npix == 0) return SR_OK;
  head1x1_bwd_kernel<<<grid_for(npix, 64, 148 * 8), 128, 0, as_stream(stream)>>>(
      x, reinterpret_cast<const __nv_bfloat16*>(act_bf16), g_f32, reinterpret_cast<const __nv_bfloat16*>(g_bf16),
      npix, dw, db);
  return check_launch("head1x1_bwd_kernel");
}
