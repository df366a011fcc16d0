// Implicit-GEMM convolution for sm_100a: TMA halo staging -> tcgen05.mma (TMEM accumulators) ->
// fused epilogue.  Replaces every Conv2D of the DifvdsrDouble stack
// (reference: models.py:1177-1199, 1231-1270; Keras Conv2D = cross-correlation, SAME, stride 1).
//
// Formulation ("flat halo"): a column segment of the image (width BW) is staged in shared memory
// together with its zero halo as a dense strip of NR rows with pitch PWs = BW + 2p pixels (TMA box
// load; out-of-bounds pixels are zero-filled by the TMA unit, which IS the SAME padding).  In the
// strip every filter tap (dy,dx) is a constant pixel shift dy*PWs + dx, so the A operand of tap
// (dy,dx) is the same shared-memory strip addressed through a UMMA descriptor whose start address
// is shifted by that many pixel rows: the strip is loaded once per 32-channel chunk and reused
// by all k*k taps.  One CTA tile = T = NACC*128 consecutive flat positions (NACC accumulators of
// 128x128 fp32 in TMEM), so every weight stage (tap, 32 channels, 128 couts = 8 KB) feeds
// NACC*2 MMAs.  Positions that fall on halo columns are computed and discarded (4 of 100).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/sr100.h"

namespace sr {

constexpr int kConvThreads = 256;  // warp0: weight TMA, warp1: MMA, warp2: TMEM alloc, warp3: A TMA, warps4-7: epilogue
constexpr int kCin = 128;          // every tensor-core conv of the stack has 128 input channels
constexpr int kChunk = 32;         // channels per K chunk (64-byte rows)
constexpr int kNumChunks = kCin / kChunk;
constexpr int kRowBytes = kChunk * 2;  // one pixel row of a K chunk in shared memory: 64 bytes (TMA / UMMA 64B swizzle)

// Operand precision.  bf16: 32 channels per 64-byte row, 4 K chunks, kind::f16 MMAs of K = 16.
// tf32: the activations and packed weights are fp32 words (rounded to tf32 by their producers), 16 channels
// per 64-byte row, 8 K chunks, kind::tf32 MMAs of K = 8 -- byte for byte the same strips, weight stages and
// descriptors, twice as many of them.
template <bool TF32>
struct PrecCfg {
  static constexpr int kElems = TF32 ? 16 : 32;   // channels per K chunk
  static constexpr int kChunks = kCin / kElems;
};

enum ConvAMode : int {
  kAModeSwizzle64 = 0,   // A strip [pixel][32ch] with TMA/UMMA 64B swizzle, tap = start-address shift
  kAModeInterleave = 1,  // A strip [8ch group][pixel][8ch] no swizzle (SBO=128B, LBO=strip bytes/4)
};

struct ConvKernelParams {
  int nsrc;          // 1 or 2 K-concatenated sources (second half of a 5/3 block)
  int ksize[2];
  int nch[2];        // K chunks per source actually multiplied (sr_conv_desc.cin_valid; all of them by default)
  int H, W, NB;     // tensor dims (addressing, TMA bounds)
  int Hc, Wc;       // compute extents (<= H, W): pixels outside are neither produced nor stored
  int p;             // geometry halo = max (k-1)/2 over sources
  int BW, nseg, PWs, NR;
  int tiles_per_seg, total_tiles;
  int a_bytes;       // bytes of one A strip buffer (1024-aligned)
  int num_abuf;      // strip buffers in flight (2; up to 4 in the pair kernel for launches of less than one wave)
  int num_wstages;
  int f_len;         // number of useful flat positions per segment
  // epilogue: out = act(alpha * (acc + bias) + beta * res)
  const float* bias;
  float alpha, beta;
  int relu;          // 1: ReLU, 2: LeakyReLU(neg_slope) (generic epilogue only)
  float neg_slope;
  const float* res_f32;
  const __nv_bfloat16* res_bf16;
  __nv_bfloat16* out_bf16;
  float* out_f32;
  float* colsum;     // training: colsum[ch] += colsum_scale * sum over stored pixels of the bf16 output (or null)
  float colsum_scale;
  float* out_tf32;   // tf32 mode: the output again, rounded to tf32 (round-to-nearest) = the next conv's operand
  int cout;          // real output channels (128; 3 for the tail conv)
  // dgrad-time ReLU mask: if non-null, out *= (mask > 0)
  const __nv_bfloat16* relu_mask_bf16;
  float mask_slope;  // factor where the mask is <= 0 (0 = ReLU backward; LeakyReLU backward: generic epilogue only)
  // cout <= 16 path only: image n is written to slot out_index[n] of a tensor of out_H x out_W pixel images
  const int* out_index;
  int out_H, out_W;
  // cout <= 16 path only: fused quantise + stitch (sr_conv_desc.stitch_*); tiles indexed like out_index slots
  const sr_stitch_tile* stitch_tiles;
  unsigned char* stitch_u8;
  float stitch_mul;
  // sub-pixel layers: r > 0 stores channel ch of pixel (y,x) at the depth-to-space position (order as sr_depth_to_space)
  int shuffle_r, shuffle_order, shuffle_C;
  // development switches of the pair kernel (SR100_CONV_DBG, compiled in only with -DSR_DEV_SWITCHES; results are WRONG with either set -- timing only):
  // 1 = no weight TMA after the first pass over the stage ring, 2 = no activation-strip TMA after the first two
  int dbg;
  // development build only: per-CTA clock64 stamps of the pair kernel's phases (sr_dev_set_timeline), 16 per CTA
  unsigned long long* timeline;
};

}  // namespace sr
