/* sr100 -- C ABI of the B200-native x4 super-resolution hot path.
 *
 * The reference (diacaf/image-enhance-keras) has no FFI layer: its hot path is Keras/TensorFlow
 * ops reached from Python (models.py) plus numpy helpers (img_utils.py, PSNR.py, scorpath.py).
 * This header is the boundary a binding for that path would use: every entry point names the
 * reference operation it replaces (file:line in /root/reference).  All pointers are DEVICE
 * pointers unless the name ends in _host; `stream` is a cudaStream_t passed as void*.
 * Every function returns 0 (SR_OK) or a negative SR_ERR_* code; sr_last_error_string() gives the
 * thread-local message.  Functions never allocate or free caller memory and are asynchronous on
 * `stream` unless stated otherwise.
 */
#ifndef SR100_H_
#define SR100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  SR_OK = 0,
  SR_ERR_INVALID = -1,     /* bad argument (shape, null pointer, patch larger than image ...) */
  SR_ERR_UNSUPPORTED = -2, /* shape/dtype outside what the sm_100a kernels implement */
  SR_ERR_CUDA = -3,        /* CUDA runtime / driver error; message holds cudaGetErrorString */
  SR_ERR_NOMEM = -4
};

const char* sr_last_error_string(void);
int sr_version(void);
/* Returns 1 when the current device is compute capability 10.x (sm_100a cubins can run). */
int sr_device_supported(void);
/* 1 when the library was built with -DSR_DEV_SWITCHES (development build: environment variables such as
 * SR100_CONV_DBG can make kernels skip operand loads, i.e. give wrong results for timing experiments); the
 * default build returns 0 and reads no environment variable that can change a result. */
int sr_dev_switches(void);
/* Programmatic dependent launch of the tensor-core kernels (default 1): a conv / wgrad launch may become resident
 * while its predecessor in the stream drains (its set-up and the launch latency leave the critical path; the kernels
 * wait for the predecessor before their first global access).  0: plain stream-ordered launches; same results.
 * Applies to launches and graph captures made afterwards; returns the previous setting. */
int sr_set_pdl(int enabled);
/* Development build only (SR_ERR_UNSUPPORTED otherwise): conv plans created after this call stamp per-CTA phase clocks
 * (16 x uint64 per CTA) into `buf` -- the intra-kernel timeline behind profiles/r02_probe_timeline*.json. */
int sr_dev_set_timeline(void* buf);
/* sizeof of the ABI structs as compiled into the library, for bindings to check their own declarations against:
 * 0 sr_conv_desc, 1 sr_conv_plan_info_t, 2 sr_pack_item, 3 sr_wgrad_desc, 4 sr_wgrad_plan_info_t,
 * 5 sr_score_result, 6 sr_model_config, 7 sr_forward_desc, 8 sr_train_desc, 9 sr_model_run_info,
 * 10 sr_stitch_tile, 11 sr_score_item; 0 for anything else. */
size_t sr_abi_struct_size(int which);

/* ------------------------------------------------------------------------------------------
 * Convolution (tensor cores).  Replaces keras Conv2D(padding='same') of the DifvdsrDouble stack:
 * models.py:1177-1199 (stack), :1231-1245 (_residual_block_light), :1248-1270
 * (_residual_block_light53), with the scalar_mul/Add lambdas (models.py:977-986) fused:
 *     out = act( alpha * (sum_s conv(in[s], w[s]) + bias) + beta * res )
 * Activations are NHWC bf16 with 128 channels; accumulation is fp32 in TMEM.
 * ------------------------------------------------------------------------------------------ */
/* One patch of a fused tail-conv stitch (sr_conv_desc.stitch_tiles): where the patch sits in its image and which of its
 * pixels it OWNS under the last-writer-wins overwrite order of img_utils.rebuild_from_patches_Step
 * (img_utils.py:700-722: ownership passes from tile i-1 to tile i at scale*step*i + 8). */
typedef struct sr_stitch_tile {
  long long img_offset;  /* byte offset of the patch's uint8 image [img_h, img_w, 3] inside stitch_u8 */
  int img_h, img_w;      /* image size in pixels; writes outside are dropped (the crop to 4H x 4W, models.py:412) */
  int y0, x0;            /* image coordinates of patch pixel (0, 0) (x0 may be negative for a column strip) */
  int oy0, oy1, ox0, ox1;/* owned patch-local pixel range [oy0, oy1) x [ox0, ox1) */
} sr_stitch_tile;

typedef struct sr_conv_desc {
  int nsrc;              /* 1, or 2 = two convolutions accumulated into one output (5/3 block tail) */
  const void* in[2];     /* bf16 [NB,H,W,128] */
  const void* wpacked[2];/* from sr_pack_conv_weights */
  int ksize[2];          /* 1, 3, 5 (7) */
  int NB, H, W;
  int cin;               /* must be 128 */
  int cout;              /* 128, or <= 16 (tail conv: 3) */
  const float* bias;     /* [cout] fp32 or NULL; for nsrc == 2 the caller passes bias0 + bias1 */
  float alpha, beta;
  int relu;              /* 1: ReLU after the residual add; 2: LeakyReLU(leaky_slope), keras LeakyReLU of
                          * Difvdsr4._residual_block_light0 (models.py:1134) / Difvdsr._residual_block (:1351) */
  const float* res_f32;  /* optional residual, fp32 [NB,H,W,cout] */
  const void* res_bf16;  /* optional residual, bf16 (used when res_f32 is NULL) */
  void* out_bf16;        /* optional bf16 output [NB,H,W,cout] */
  float* out_f32;        /* optional fp32 output [NB,H,W,cout] */
  const void* relu_mask_bf16; /* optional (backward): out = 0 where mask <= 0 */
  int a_mode;            /* 0: 64B-swizzled strip, 1: interleaved no-swizzle strip */
  int nacc;              /* 4 (default): 512 positions / tile; 2: 256 positions, double-buffered TMEM */
  int pair;              /* 1: CTA-pair kernel (tcgen05 cta_group::2, M = 256 over two image x column-segment columns); NB >= 2 */
  /* cout <= 16 only (the tail conv): scatter image n of the output into slot out_index[n] (device
   * int32[NB]) of a tensor whose images are out_h x out_w pixels (>= H x W): the cropped HR stage
   * writes its H x W result into the top-left corner of the full 384 x 384 patch slot. NULL: dense. */
  const int* out_index;
  int out_h, out_w;
  /* Sub-pixel layers (keras_subpixel.Subpixel, advanced.SubPixelUpscaling / SubpixelConv2D): r > 0
   * fuses the depth-to-space shuffle into the epilogue's store address; cout = r*r*C may then be
   * any value <= 128 (weights packed with sr_pack_conv_weights(cout)), output fp32
   * [NB, H*r, W*r, C] only; order as in sr_depth_to_space. */
  int shuffle_r, shuffle_order;
  /* Compute extents: > 0 restricts the launch to the top-left comp_h x comp_w corner of every image
   * (outputs outside are not written; inputs outside are read as they are, only the true image
   * border is zero-padded).  Lets the tiled path shrink the last LR layers to the region whose
   * receptive field can still reach a surviving pixel.  0: the whole H x W image. */
  int comp_h, comp_w;
  float leaky_slope;     /* negative-side slope when relu == 2 */
  /* Operand precision.  0: bf16 (above).  1: tf32 -- the "tf32 option" of the conv stack for users who want
   * the fp32 graph of the reference (Keras/TF Conv2D is fp32) to ~1e-4: in[] are fp32 [NB,H,W,128] whose values
   * are already rounded to tf32 (out_tf32 of the producing launch, sr_round_tf32), wpacked[] come from
   * sr_pack_conv_weights_tf32, MMAs are tcgen05 kind::tf32 with fp32 accumulation, the epilogue is fp32
   * throughout.  Outputs: out_f32 (unrounded: the residual stream) and/or out_tf32; residual: res_f32 only.
   * bf16 tensors, relu_mask, shuffle_r, LeakyReLU and a_mode 1 are not available in this mode. */
  int precision;
  float* out_tf32;       /* precision 1, cout == 128: the output rounded to tf32 (round to nearest) */
  /* Training: colsum_f32[c] += colsum_scale * sum over all pixels of the (bf16-rounded) output -- the bias gradient
   * of the layer whose output gradient this input-gradient launch writes (what sr_colsum_bf16 would compute in a
   * second pass over the tensor).  fp32 [128], atomically accumulated; needs out_bf16, cout == 128, nacc == 2 and
   * exactly one of res_f32 / res_bf16 / relu_mask_bf16. */
  float* colsum_f32;
  float colsum_scale;
  /* cout <= 16 only (the tail conv): quantise + stitch fused into the epilogue.  Patch n of the
   * launch (slot out_index[n] when out_index is given) is described by stitch_tiles[slot] (DEVICE array); every
   * output pixel the patch owns is written as uint8 = trunc(clip(value * stitch_mul, 0, 255)) (models.py:351, 391)
   * straight into its image inside stitch_u8 -- the fp32 patch tensor (out_f32, may then be NULL) and the separate
   * sr_patch_stitch pass disappear.  Bit-identical to conv -> sr_patch_stitch. */
  const sr_stitch_tile* stitch_tiles;
  uint8_t* stitch_u8;
  float stitch_mul;
  /* cin_valid[s] > 0: only the first cin_valid[s] input channels of source s are real (a multiple of 32; 16 for
   * tf32), the rest of its 128 are zero padding that the launch does not load or multiply -- the second plane of a
   * 192-channel tensor (Difvdsr, models.py:1274-1357) costs K = 64 instead of 128.  0: all 128. */
  int cin_valid[2];
  /* With relu_mask_bf16: the factor applied where the mask is <= 0 (0: ReLU backward; alpha: the backward of
   * LeakyReLU(alpha), models.py:1024-1032 / 1345-1350, whose output has the sign of its input). */
  float relu_mask_slope;
} sr_conv_desc;

typedef struct sr_conv_plan sr_conv_plan;

typedef struct sr_conv_plan_info_t {
  double flops;          /* algorithmic FLOPs (2*MAC) per run */
  double mma_efficiency; /* useful / issued MMA rows */
  int total_tiles, grid, smem_bytes, seg_width, nseg, strip_rows, num_wstages, tile_positions;
} sr_conv_plan_info_t;

/* Synchronous, host side only: picks tile geometry and encodes the TMA descriptors.  The plan is
 * bound to the pointers in the descriptor. */
int sr_conv_plan_create(const sr_conv_desc* desc, sr_conv_plan** plan);
int sr_conv_plan_run(sr_conv_plan* plan, void* stream);
void sr_conv_plan_destroy(sr_conv_plan* plan);
int sr_conv_plan_info(const sr_conv_plan* plan, sr_conv_plan_info_t* info);

/* A SEQUENCE of 128 -> 128 bf16 convolutions in ONE persistent launch, for inputs so small that a per-layer launch is
 * mostly launch gap, setup and teardown (one 128 x 128 patch through model.predict, models.py:342 / 165-182: BASELINE
 * config 1).  descs[i] as for sr_conv_plan_create (plain bf16-out epilogue, or fp32 residual + fp32 / bf16 outputs);
 * phase[i] (HOST array, non-decreasing from 0) groups convolutions that do not depend on each other -- the two branch
 * heads of a 5/3 block read the same tensor -- and phases are separated by a grid-wide barrier inside the kernel:
 * every output of phase p is visible before any input of phase p + 1 is read.  Results are bit-identical to running
 * the plans one by one.  The chain owns a small device buffer (tensor maps + barrier counter).
 * Measured on B200 (tools/probe_chain.py, profiles/r02_probe_chain.txt): per phase the barrier costs ~4 us and the
 * strip burst + epilogue of a one-tile phase stay exposed, so the chain is SLOWER than per-layer launches replayed from
 * a CUDA graph (config 1: 2.33 vs 2.06 ms); kept as an entry point and as the record of that experiment. */
typedef struct sr_conv_chain sr_conv_chain;
int sr_conv_chain_create(const sr_conv_desc* descs, const int* phase, int n, sr_conv_chain** chain);
int sr_conv_chain_run(sr_conv_chain* chain, void* stream);
void sr_conv_chain_destroy(sr_conv_chain* chain);
/* flops, grid, smem_bytes, num_wstages, total_tiles; strip_rows = number of phases, nseg = number of convolutions */
int sr_conv_chain_info(const sr_conv_chain* chain, sr_conv_plan_info_t* info);

/* Repack Keras HWIO fp32 weights [k,k,cin=128,cout] (device) into the kernel's K-chunked bf16
 * layout [cin/32][k*k][cout_pad][32]; cout_pad = 16 for cout <= 16, else 128 (zero rows beyond cout).  transpose_flip = 1 produces the
 * weights of the input-gradient convolution (180-degree rotation, cin<->cout): always a
 * 128 -> 128 layout (destination size sr_packed_weight_bytes(ksize, 128)); for cout < 128 the
 * missing reduction rows are zero, so the gradient tensor may carry anything in channels >= cout.
 * sr_packed_weight_bytes gives the destination size. */
size_t sr_packed_weight_bytes(int ksize, int cout);
/* Many layers in one launch (after an optimizer step every layer is repacked, forward and transposed).
 * items / starts are DEVICE arrays: starts[i] = sum of the packed ELEMENT counts
 * (sr_packed_weight_bytes / 2) of items 0..i-1, starts[n_items] = total_elems. */
typedef struct sr_pack_item {
  const float* hwio;     /* [ksize,ksize,128,cout] fp32 */
  void* dst;             /* packed bf16 destination */
  int ksize, cout, transpose_flip, pad_;
} sr_pack_item;
int sr_pack_conv_weights_batched(const sr_pack_item* items_dev, const unsigned long long* starts_dev,
                                 int n_items, size_t total_elems, void* stream);
int sr_pack_conv_weights(const float* hwio, int ksize, int cout, int transpose_flip, void* dst,
                         void* stream);
/* The same repack for sr_conv_desc.precision == 1: fp32 words rounded to tf32 (round to nearest), layout
 * [cin/16][k*k][cout_pad][16] (64-byte rows like the bf16 layout, twice as many K chunks). */
size_t sr_packed_weight_bytes_tf32(int ksize, int cout);
int sr_pack_conv_weights_tf32(const float* hwio, int ksize, int cout, void* dst, void* stream);
/* out[i] = in[i] rounded to tf32 (cvt.rna): makes an fp32 tensor a tf32 conv operand (the tensor core would
 * otherwise truncate the low 13 mantissa bits, a biased error).  in == out is allowed. */
int sr_round_tf32(const float* in, size_t n, float* out, void* stream);

/* Plain CUDA-core convolution, fp32 accumulate, any cin/cout, SAME or VALID, NHWC.  Used for
 * layers outside the 128-channel stack (Subpixel(Conv2D), keras_subpixel.py:28-62) and as the
 * on-device cross-check of the tensor-core kernel.  in_is_bf16/w_round_bf16 reproduce the
 * tensor-core kernel's operand rounding.  shuffle_r > 0 fuses the depth-to-space store, see
 * sr_depth_to_space for the orderings. */
int sr_conv2d_direct(const void* in, int in_is_bf16, const float* hwio, int w_round_bf16,
                     const float* bias, int NB, int H, int W, int cin, int cout, int ksize,
                     int same_padding, int relu, int shuffle_r, int shuffle_order, float* out,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * First layer.  Replaces Convolution2D(128,(1,1),relu,name='level1') (models.py:1177) applied to
 * the float32 /255 patch stack (models.py:336).  in: fp32 [NPIX,3] in [0,1]; w: fp32 [3,128]
 * (HWIO of the 1x1 kernel); out: bf16 [NPIX,128] and optionally fp32 [NPIX,128].
 * ------------------------------------------------------------------------------------------ */
int sr_head1x1_fwd(const float* in, const float* w, const float* bias, size_t npix, void* out_bf16,
                   float* out_f32, void* stream);

/* ------------------------------------------------------------------------------------------
 * Bilinear x4, TensorFlow-1 legacy sampling (align_corners=False, no half-pixel centres).
 * Replaces Lambda(resizeX4bil) = tf.image.resize_bilinear (models.py:1193, 1392-1399).
 * in: [NB,H,W,C] -> out: [NB,4H,4W,C]; C % 8 == 0.  Either in/out dtype may be fp32 or bf16.
 * ------------------------------------------------------------------------------------------ */
int sr_bilinear4_fwd(const void* in, int in_is_bf16, int NB, int H, int W, int C, void* out_bf16,
                     float* out_f32, void* stream);
/* Same sampling, restricted to the top-left out_h x out_w corner of the x4 output (multiples of 4)
 * and with an optional gather of the source images: out[n] = resize(in[src_index[n]])[:out_h,:out_w].
 * Used by the tiled inference path: of a 384x384 tile output only [8,264) per axis survives the
 * stitch (img_utils.py:700-722) and the HR stage (2 blocks + tail) has a 7-pixel receptive-field
 * radius, so 272 x 272 of the upsampled tile is all the HR stage ever needs. */
int sr_bilinear4_crop_fwd(const void* in, int in_is_bf16, const int* src_index, int n_out, int H,
                          int W, int C, int out_h, int out_w, void* out_bf16, float* out_f32,
                          void* stream);
/* Bilinear x2 with the same TF1 legacy sampling (src = dst * 0.5): Lambda(resize2bil) =
 * tf.image.resize_bilinear(x, [2h,2w]) of Difvdsr4 (models.py:932-940, applied at :1046 and :1053). */
int sr_bilinear2_fwd(const void* in, int in_is_bf16, int NB, int H, int W, int C, void* out_bf16,
                     float* out_f32, void* stream);
/* Adjoint of the above: gin[NB,H,W,C] = sum over the HR samples each LR pixel contributed to. */
int sr_bilinear4_bwd(const float* gout, int NB, int H, int W, int C, float* gin, void* stream);
/* The adjoint of sr_bilinear2_fwd (Lambda(resize2bil) of Difvdsr4, models.py:932-940, in training): gout fp32
 * [NB,2H,2W,C] -> gin fp32 [NB,H,W,C]. */
int sr_bilinear2_bwd(const float* gout, int NB, int H, int W, int C, float* gin, void* stream);

/* ------------------------------------------------------------------------------------------
 * Patch tiling.  Replaces img_utils.extract_patches_Step (img_utils.py:601-676) and
 * img_utils.rebuild_from_patches_Step (img_utils.py:692-724) as used by
 * BaseSuperResolutionModel.upscaleStepPatch (models.py:184-415).
 * ------------------------------------------------------------------------------------------ */
/* Number of patch positions per axis: |{x : 0 <= x < dim - patch, x % step == 0}| (img_utils.py:622,629). */
int sr_patch_count(int dim, int patch, int step);
/* Zero-padded canvas size of upscaleStepPatch (models.py:225-256): +patch border, then both
 * dims bumped to int(x/step+1)*step when either is not a multiple of step. */
int sr_canvas_size(int h, int w, int patch, int step, int* canvas_h, int* canvas_w);
/* Gather: uint8 image [h,w,3] placed at the top-left of a zero canvas [ch,cw] -> patches
 * [cnt_w*cnt_h, ph, pw, 3], column-major patch order n = wi*cnt_h + hi.  The pixel value is divided
 * by `divisor` in fp32 (1 -> the reference's 0..255 patches; 255 -> the /255. of models.py:336). */
int sr_patch_gather_u8(const uint8_t* img, int h, int w, int canvas_h, int canvas_w, int ph, int pw,
                       int step, float divisor, float* out_f32, void* stream);
/* The same gather for n_img images of one shape in ONE launch (a batch of equal-sized inputs, BASELINE config 3):
 * image m starts img_stride bytes after image m-1, its patches follow those of image m-1 in out_f32. */
int sr_patch_gather_u8_batched(const uint8_t* imgs, int n_img, size_t img_stride, int h, int w,
                               int canvas_h, int canvas_w, int ph, int pw, int step, float divisor,
                               float* out_f32, void* stream);
/* Generic gather from a float64/float32 canvas (API parity with the numpy function). */
int sr_patch_gather_f32(const float* canvas, int canvas_h, int canvas_w, int ph, int pw, int step,
                        float* out_f32, void* stream);
/* Stitch: patches [N, ph*scale, pw*scale, 3] fp32 -> canvas [ch*scale, cw*scale, 3]; closed-form
 * last-writer-wins ownership with the 8-px border crop (img_utils.py:700-722).  mul scales the
 * value (255 for models.py:351).  out_f32 and/or out_u8 (np.clip(0,255).astype('uint8'),
 * truncation, models.py:391) may be given. */
int sr_patch_stitch(const float* patches, int cnt_h, int cnt_w, int ph, int pw, int step, int scale,
                    int canvas_h, int canvas_w, float mul, float* out_f32, uint8_t* out_u8,
                    void* stream);

/* The stitch of ONE SHARD of a tile-sharded image (the exchange of img_utils.py:692-724 when the
 * tiles of one image ran on several GPUs): patches holds tiles [tile_lo, tile_hi) of the column-major tile index
 * only; out_u8 is the uint8 strip [canvas_h*scale, strip_w, 3] of output columns [x0, x0 + strip_w) -- pixels owned
 * by a tile of the range carry the stitched value (x mul, clip, truncate), all others 0.  The strips of all
 * shards OR-ed together are bit-identical to sr_patch_stitch over all tiles. */
int sr_patch_stitch_range(const float* patches, int cnt_h, int cnt_w, int ph, int pw, int step, int scale,
                          int canvas_h, int tile_lo, int tile_hi, int x0, int strip_w, float mul,
                          uint8_t* out_u8, void* stream);

/* Minibatch assembly from an HBM-resident dataset.  Replaces the per-batch file decode of
 * img_utils.image_generator (img_utils.py:341-372): data = N decoded uint8 images of item_bytes bytes each
 * (item_bytes % 4 == 0), index = device int64[n] chosen by img_utils._index_generator;
 * out[i] = float32(data[index[i]]) / divisor  (255 -> the reference's astype('float32') / 255.). */
int sr_batch_gather_u8(const uint8_t* data, size_t item_bytes, size_t n_items, const long long* index,
                       int n, float divisor, float* out_f32, void* stream);

/* ------------------------------------------------------------------------------------------
 * Alternative tilers: BaseSuperResolutionModel.upscalePatch (models.py:419-604) and
 * upscale(mode='patch') (models.py:645-680, 758-790).
 * ------------------------------------------------------------------------------------------ */
/* Patches n in [n0, n1) of the grid n = a*cnt_w + b at (a*step, b*step) of a uint8 image [H,W,3], each p x p
 * (p % 4 == 0), shrunk x4 like scipy.misc.imresize(patch, (p/4, p/4), interp='bicubic') (models.py:490, 672):
 * stretch != 0 applies scipy.misc.bytescale first (the reference's float64 patches of upscalePatch are
 * contrast-stretched to [0,255] per patch; uint8 patches of upscale() are not), then Pillow's 8-bit bicubic
 * (horizontal pass, uint8 intermediate, vertical pass) with the fixed-point coefficient table (bounds
 * int32[p/4][2] = (first tap, tap count), kk int32[p/4][ksize]) the host computes as Resample.c does.
 * out[n - n0] = float32 [p/4, p/4, 3] = value / divisor. */
int sr_patch_down4_u8(const uint8_t* img, int H, int W, int p, int step, int cnt_h, int cnt_w,
                      long long n0, long long n1, int stretch, const int* bounds, const int* kk, int ksize,
                      float divisor, float* out_f32, void* stream);
/* Averaging stitch of img_utils.reconstruct_from_patches_2dlocal (img_utils.py:442-511; step 4, pad 4: interior
 * patches contribute rows/columns [pad, P-pad) only; a patch is interior when its grid indices are > 0 and differ
 * from edge_a / edge_b, the indices sitting on the last dense position, -1 if the grid does not reach it) and of sklearn reconstruct_from_patches_2d behind
 * img_utils.combine_patches (img_utils.py:189-193; step 1, pad 0).  patches: float32 [(a1-a0)*cnt_w, P, P, 3] =
 * grid rows a0..a1-1; sum (float64 [out_h,out_w,3]) and count (int32 [out_h,out_w]) accumulate across calls made
 * in increasing a0 (zero them first): the float64 additions then happen in the reference's (i, j) order.
 * finalize: img = sum / count - or, closed_P > 0, sum / (min(Y+1,P,H-Y) * min(X+1,P,W-X)), sklearn's closed-form
 * overlap count (it exceeds the true count on images smaller than 2P-1; the reference divides by it regardless);
 * out_u8 = np.clip(img, 0, 255).astype('uint8') (models.py:575, 791). */
int sr_patch_average_accumulate(const float* patches, int P, int step, int pad, int cnt_h, int cnt_w,
                                int a0, int a1, int edge_a, int edge_b, float mul, int out_h, int out_w,
                                double* sum, int* count, void* stream);
int sr_patch_average_finalize(const double* sum, const int* count, int out_h, int out_w, int closed_P,
                              double* out_f64, uint8_t* out_u8, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dataset preparation.  Replaces the per-image body of img_utils.transform_images (img_utils.py:70-117):
 * imresize to 256 x 256 (Pillow BILINEAR), scipy.misc.imfilter 'sharpen' (Pillow SHARPEN), 256 sub-images,
 * per sub-image bytescale (imsave), scipy.ndimage.gaussian_filter(sigma 0.5) over all three axes, bytescale +
 * Pillow BICUBIC down (and up again).  Everything bit-exact against those libraries.
 * ------------------------------------------------------------------------------------------ */
/* Pillow's 8-bit two-pass resize (Resample.c) of NB uint8 [H,W,3] images to [out_h,out_w,3]: horizontal pass into
 * tmp (uint8 [NB,H,out_w,3]; may be NULL when only one axis changes), then vertical.  bounds_* int32 [out][2] =
 * (first tap, tap count), kk_* int32 [out][ksize] = filter taps with 22 fractional bits, computed on the host the
 * way precompute_coeffs / normalize_coeffs_8bpc do (any filter: BILINEAR, BICUBIC ...). */
int sr_resize_u8(const uint8_t* src, int NB, int H, int W, int out_h, int out_w, const int* bounds_x,
                 const int* kk_x, int ksize_x, const int* bounds_y, const int* kk_y, int ksize_y, uint8_t* tmp,
                 uint8_t* dst, void* stream);
/* PIL ImageFilter.SHARPEN = scipy.misc.imfilter(img, 'sharpen') (img_utils.py:75); dst != src. */
int sr_sharpen3x3_u8(const uint8_t* src, int NB, int H, int W, uint8_t* dst, void* stream);
/* Sub-images n = 0..n_patches-1 of a uint8 [H,W,3] image, P x P at (row, col) = pos[n] (device int32 [n][2]):
 * y_u8[n] = bytescale(float64 sample) (what imsave writes, img_utils.py:100); g_u8[n] = bytescale(gaussian_filter(
 * sample, sigma)) with weights[0..radius] = the normalised kernel from its centre outwards (float64, host-computed
 * like scipy.ndimage._gaussian_kernel1d), reflect boundaries, all three axes.  P <= 64. */
int sr_dataprep_patches(const uint8_t* img, int H, int W, const int* pos, int n_patches, int P,
                        const double* weights, int radius, uint8_t* y_u8, uint8_t* g_u8, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sub-pixel shuffles.  order 0: keras_subpixel.Subpixel._phase_shift (keras_subpixel.py:64-84)
 * and advanced.depth_to_scale_tf (advanced.py:104-129): ch = c*r*r + (X%r)*r + (Y%r);
 * order 1: advanced.depth_to_scale_th (advanced.py:87-100): ch = c*r*r + (Y%r)*r + (X%r);
 * order 2: tf.depth_to_space used by advanced.SubpixelConv2D (advanced.py:173-199):
 *          ch = ((Y%r)*r + (X%r))*C + c.   in: fp32 NHWC [NB,H,W,C*r*r] -> out [NB,H*r,W*r,C].
 * ------------------------------------------------------------------------------------------ */
int sr_depth_to_space(const float* in, int NB, int H, int W, int C, int r, int order, float* out,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Scoring.  Replaces scorpath.py:174-228: crop_border (scorpath.py:67-70), skimage rgb2ycbcr Y
 * (scorpath.py:26-31), PSNR.psnrNITRE / PSNRTorch (PSNR.py:24-32, 54-84) and skimage
 * compare_ssim (7x7 uniform window, sample covariance, data_range 255) on Y and on RGB.
 * ------------------------------------------------------------------------------------------ */
/* Y = 16 + (65.481 R + 128.553 G + 24.966 B)/255 in fp64, from uint8 RGB. */
int sr_rgb2y_u8(const uint8_t* rgb, size_t npix, double* y, void* stream);
typedef struct sr_score_result {
  double sum_sq_y;   /* sum over cropped pixels of (Y1 - Y2)^2, Y in [16,235] */
  double ssim_y_sum; /* sum of the SSIM map (valid 7x7 windows) on Y */
  double ssim_rgb_sum[3];
  int64_t n_pix;     /* cropped pixel count */
  int64_t n_win;     /* number of valid windows */
  /* internal: integer accumulators (2^-40 fixed-point SSIM sums of Y, R, G, B; low / high halves of the exact
   * integer squared error in units of 1/255000^2) and the count of finished blocks.  All cross-block sums are
   * integer atomics, so the public fields above -- written by the last block -- are bit-reproducible. */
  uint64_t acc[6];
  uint64_t ticket;
} sr_score_result;
/* a, b: uint8 [h,w,3] (same shape), crop = border removed from every side first.  `result` is a
 * device pointer to one sr_score_result (112 bytes) which must be zeroed by the caller before the call. */
int sr_score_pair_u8(const uint8_t* a, const uint8_t* b, int h, int w, int crop,
                     sr_score_result* result, void* stream);
/* The same for n pairs (any mix of shapes) in one launch per 32 pairs -- the loop over a directory of
 * scorpath.py:92-228 with one launch and one read-back instead of one per pair.  `items` is a HOST array (device
 * image pointers inside), `results` a device array of n zeroed sr_score_result. */
typedef struct sr_score_item {
  const uint8_t* a;
  const uint8_t* b;
  int h, w;
} sr_score_item;
int sr_score_batch_u8(const sr_score_item* items, int n, int crop, sr_score_result* results, void* stream);

/* *out += sum (a[i]-b[i])^2 in fp64 (device accumulator, zero it first): the reduction inside
 * PSNR.psnrVDSR / PSNRTorch / psnrSVLAB / psnrNITRE (PSNR.py:7-84) and models.psnr* (models.py:71-90). */
int sr_sum_sq_diff_f64(const double* a, const double* b, size_t n, double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training step pieces (models.py:131-157 fit, :1212-1213 compile: mse + Adam(1e-4, 0.9)).
 * The input gradient of a conv is sr_conv_plan_* with transpose_flip weights (+ relu_mask_bf16).
 * ------------------------------------------------------------------------------------------ */
/* Filter gradient of a 128 -> 128 Conv2D(padding='same') on the tensor cores:
 *   dw[ky][kx][ci][co] (HWIO fp32) = (accumulate ? dw : 0) + scale * sum_{n,y,x} x[n,y+ky-p,x+kx-p,ci] * g[n,y,x,co]
 * x, g: bf16 NHWC [NB,H,W,128].  workspace: device scratch of sr_wgrad_workspace_bytes() bytes
 * (per-CTA fp32 partials, summed in a fixed order: results are run-to-run deterministic). */
typedef struct sr_wgrad_desc {
  const void* x_bf16;
  const void* g_bf16;
  int NB, H, W;
  int ksize;             /* 1, 3 or 5 */
  float scale;
  int accumulate;
  float* dw_hwio;        /* [ksize,ksize,128,128] fp32 */
  void* workspace;
  size_t workspace_bytes;
} sr_wgrad_desc;
typedef struct sr_wgrad_plan sr_wgrad_plan;
typedef struct sr_wgrad_plan_info_t {
  double flops;
  int grid, smem_bytes, seg_width, nseg, ring_rows, g_slots, tap_groups, rows_per_unit, images_per_row;
} sr_wgrad_plan_info_t;
size_t sr_wgrad_workspace_bytes(void);
int sr_wgrad_plan_create(const sr_wgrad_desc* desc, sr_wgrad_plan** plan);
int sr_wgrad_plan_run(sr_wgrad_plan* plan, void* stream);
void sr_wgrad_plan_destroy(sr_wgrad_plan* plan);
int sr_wgrad_plan_info(const sr_wgrad_plan* plan, sr_wgrad_plan_info_t* info);

/* Loss gradient at the tail: loss_sum += sum (pred-target)^2 (fp64, device); g128[pix][c] =
 * 2*(pred-target)/n_total where pred > 0 (ReLU of the last Conv2D, models.py:1199), as bf16 rows
 * of 128 channels (channels >= `channels` zero) = the operand layout of dgrad / wgrad. */
int sr_mse_tail_grad(const float* pred, const float* target, size_t npix, int channels,
                     size_t n_total, void* g128_bf16, double* loss_sum, void* stream);
/* The same loss gradient for a 3-channel 3x3 tail, laid out as the im2col of its backward: channel
 * j = (ky*3+kx)*3 + co of pixel (y,x) holds g3[y-ky+1][x-kx+1][co] (0 outside the image), channels 27..127 zero.
 * Both tail gradients then are 1x1 problems for the 128-wide tensor-core kernels:
 *   dgrad = sr_conv_plan (ksize 1) of this tensor with B[j][ci] = w[ky][kx][ci][co];
 *   wgrad = sr_wgrad_plan (ksize 1) of (tail input, this tensor): dw[ky][kx][ci][co] = D[ci][j];
 *   bias gradient: db3[co] += sum over pixels of the (bf16-rounded) gradient, if db3 is not NULL. */
int sr_mse_tail_grad_col(const float* pred, const float* target, int NB, int H, int W, size_t n_total,
                         void* a128_bf16, double* loss_sum, float* db3, void* stream);
/* Bias gradient: out[c] += scale * sum_pix g[pix][c]; g bf16 [npix,128]; out fp32 [128]. */
int sr_colsum_bf16(const void* g_bf16, size_t npix, float scale, float* out, void* stream);
/* First layer (models.py:1177) backward: g0 = g*(act>0); dw[3][128] += x^T g0; db[128] += colsum g0.
 * x fp32 [npix,3]; act bf16 [npix,128] (layer output); g fp32 or bf16 [npix,128] (one of them). */
int sr_head1x1_bwd(const float* x, const void* act_bf16, const float* g_f32, const void* g_bf16,
                   size_t npix, float* dw, float* db, void* stream);
/* loss_sum += sum (pred-target)^2 (fp64 accumulator, device); grad = 2*(pred-target)/n_total. */
int sr_mse_loss_grad(const float* pred, const float* target, size_t n, size_t n_total, float* grad,
                     double* loss_sum, void* stream);
/* Keras-2 Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps).  grad_scale is
 * applied to g first (1/world_size after an all-reduce sum). */
int sr_adam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                 float beta2, float eps, int t, float grad_scale, void* stream);
/* ReLU backward / elementwise helpers on fp32 or bf16 NHWC tensors. */
int sr_axpby_f32(const float* x, const float* y, float a, float b, size_t n, float* out,
                 void* out_bf16, void* stream);
int sr_cast_f32_to_bf16(const float* in, size_t n, void* out_bf16, void* stream);
int sr_cast_bf16_to_f32(const void* in_bf16, size_t n, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Graph-level entry points.  Replace keras Model.predict (models.py:178, 342, 541, 783) and one
 * train_on_batch of Model.fit_generator (models.py:146-157) on the DifvdsrDouble graph
 * (models.py:1159-1270): the whole launch sequence (head 1x1 -> 16 5/3 blocks -> 6 light blocks ->
 * bilinear x4 -> 2 5/3 blocks -> tail conv; its backward; Adam) lives in the library, so a binding that is
 * not Python can run the model.  An sr_model owns the packed tensor-core weights, the conv / wgrad plans
 * (TMA descriptors) per shape and the CUDA graphs that replay them; activations live in a caller-provided
 * workspace (size query + pointer, the library never allocates caller-visible tensors).
 * ------------------------------------------------------------------------------------------ */
typedef struct sr_model sr_model;

typedef struct sr_model_config {
  int precision;       /* 0: bf16 operands (default); 1: tf32 operands, fp32 tensors (inference only) */
  int stream_lr_fp32;  /* 1 (default): the residual stream of the 22 LR blocks is also kept in fp32 */
  int stream_hr_fp32;  /* 1: the HR stage too (forced for tf32); default 0 */
  int a_mode, nacc, pair; /* sr_conv_desc fields of every launch; defaults 0, 2, 1 */
  int use_graphs;      /* 1 (default): a (shape, pointers) combination runs eagerly once, then is captured into a
                        * CUDA graph that later calls replay with one submission */
  int overlap_heads;   /* 1 (default): the two branch heads of a 5/3 block (independent launches) run concurrently on
                        * an internal second stream when both grids fit on the chip together (small inputs) */
  int fused_colsum;    /* 1 (default): bias gradients ride the input-gradient launches (sr_conv_desc.colsum_f32);
                        * 0: separate sr_colsum_bf16 passes (same values, another fp32 summation order) */
  int chain_lr;        /* 1: the 60 launches of the LR stage of a SMALL input (about one 128-position tile per CTA pair
                        * and layer: one 128 x 128 patch) run as one persistent sr_conv_chain launch.  Bit-identical,
                        * measured slower than the per-layer launches under a graph: default 0 */
  int overlap_train;   /* 1: independent launches of the training step (the two branch heads of a 5/3 block, their two
                        * input-gradient launches, the wgrad pairs) are issued on two streams, so the tail wave of one
                        * persistent kernel is back-filled by the first CTAs of the other; 0 (default): one stream */
} sr_model_config;
void sr_model_default_config(sr_model_config* cfg);

/* The parameter arena: all 86 layers in Keras creation order ('level1', conv2d_1 .. conv2d_85), each as
 * kernel (HWIO fp32) then bias; offsets in floats.  name: at least 16 bytes. */
int sr_model_num_layers(void);
size_t sr_model_param_count(void);
int sr_model_layer(int index, char* name, int* ksize, int* cin, int* cout, size_t* kernel_offset,
                   size_t* bias_offset);

/* params: DEVICE fp32 [sr_model_param_count()], caller-owned, read at every sr_model_refresh (and updated in
 * place by sr_model_apply_gradients).  Creating the model packs the weights once (synchronous). */
int sr_model_create(float* params, const sr_model_config* cfg, sr_model** model);
void sr_model_destroy(sr_model* model);
/* Re-pack the tensor-core weight layouts after the caller changed params (load_weights / set_weights). */
int sr_model_refresh(sr_model* model, void* stream);

typedef struct sr_forward_desc {
  int NB, H, W;
  const float* x;       /* fp32 [NB,H,W,3] in [0,1] (the /255 patch stack, models.py:336) */
  float* out;           /* fp32 [NB,4H,4W,3] (what model.predict returns) */
  void* workspace;      /* device scratch, >= sr_model_forward_workspace_bytes() */
  size_t workspace_bytes;
  /* Optional dead-region elimination of the tiled path (img_utils.py:700-722 keeps [8,264) of a 384-pixel patch
   * axis): the patches are partitioned into n_groups classes; the patches group_index[...] of class g get only
   * the top-left group_eh[g] x group_ew[g] corner (multiples of 4) of their output computed, the rest of their
   * slot is left untouched, and the last LR layers shrink to the region that corner can see.  HOST arrays;
   * group_index holds sum(group_n) = NB patch numbers, class after class.  n_groups = 0: every patch whole. */
  int n_groups;
  const int* group_eh;
  const int* group_ew;
  const int* group_n;
  const int* group_index;
  /* Optional fused stitch (sr_conv_desc.stitch_*): DEVICE array of NB tiles, patch n -> stitch_tiles[n]; the tail
   * convs then write the owned uint8 pixels into stitch_u8 and `out` may be NULL. */
  const sr_stitch_tile* stitch_tiles;
  uint8_t* stitch_u8;
  float stitch_mul;
} sr_forward_desc;

typedef struct sr_model_run_info {
  double conv_flops;    /* algorithmic FLOPs (2*MAC) of the tensor-core launches */
  int launches;         /* kernel launches of one call */
  int conv_launches;    /* of which tensor-core conv / wgrad launches */
  int graph_replay;     /* 1 when the next call will replay a captured CUDA graph */
} sr_model_run_info;

size_t sr_model_forward_workspace_bytes(const sr_model* model, const sr_forward_desc* desc);
int sr_model_forward(sr_model* model, const sr_forward_desc* desc, void* stream);
int sr_model_forward_info(sr_model* model, const sr_forward_desc* desc, sr_model_run_info* info);
/* The same forward launch by launch with CUDA events around every kernel (synchronises `stream`): ms[i], flops[i]
 * (0 for launches that are not tensor-core convs) for i < min(*n, max_records); *n = launches.  For roofline
 * measurements (bench.py); never replays a graph. */
int sr_model_forward_timed(sr_model* model, const sr_forward_desc* desc, void* stream, float* ms,
                           double* flops, int max_records, int* n);

typedef struct sr_train_desc {
  int NB, H, W;
  const float* x;       /* fp32 [NB,H,W,3] */
  const float* y;       /* fp32 [NB,4H,4W,3] target */
  float* grads;         /* fp32 [sr_model_param_count()]: zeroed, then d(mse over this batch)/d(param) */
  double* loss_sum;     /* device fp64: zeroed, then sum of squared errors of this batch */
  float* pred;          /* optional fp32 [NB,4H,4W,3]: the forward output (NULL: kept in the workspace) */
  void* workspace;
  size_t workspace_bytes;
} sr_train_desc;

size_t sr_model_train_workspace_bytes(const sr_model* model, int NB, int H, int W);
/* Forward with saved activations, loss='mse' and the whole backward (dgrad / wgrad on the tensor cores, bias and
 * head gradients): compile(loss='mse') + the gradient half of train_on_batch (models.py:1212-1213, 146). */
int sr_model_forward_backward(sr_model* model, const sr_train_desc* desc, void* stream);
int sr_model_train_info(sr_model* model, const sr_train_desc* desc, sr_model_run_info* info);
/* Keras-2 Adam (sr_adam_step) over the whole arena with the model's own params, then sr_model_refresh.
 * grad_scale: 1/world after an all-reduce(sum) of `grads` over data-parallel ranks, else 1. */
int sr_model_apply_gradients(sr_model* model, const float* grads, float* m, float* v, int t, float lr,
                             float beta1, float beta2, float eps, float grad_scale, void* stream);
/* One single-process optimizer step: sr_model_forward_backward + sr_model_apply_gradients(grad_scale 1). */
int sr_model_train_step(sr_model* model, const sr_train_desc* desc, float* m, float* v, int t, float lr,
                        float beta1, float beta2, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * The path's one exchange step (data-parallel training): average the minibatch shards' gradients and
 * apply compile(Adam(1e-4, 0.9)) (models.py:1212-1213) to identical replicas -- as ONE kernel per rank that does
 * reduce-scatter + Adam + all-gather over NVLink peer memory (csrc/exchange.cu) instead of ncclAllReduce followed by
 * a full-arena Adam pass.  One process per GPU on one node; the gradient arena, the parameter arena and a small
 * zero-initialised signal pad of every rank are shared with CUDA IPC (sr_ipc_*; the 64-byte handles travel over any
 * side channel, e.g. torch.distributed.all_gather_object).  Rank r reduces, updates and broadcasts the shard
 * sr_exchange_shard(n, r, world); its Adam state m, v is valid on that shard only (sharded optimizer state).
 * All ranks must call sr_exchange_adam_step the same number of times; waits on peers are bounded
 * (sr_exchange_set_timeout_ms, default 30 s) and a time-out is reported by sr_exchange_status instead of a hang.
 * ------------------------------------------------------------------------------------------ */
#define SR_IPC_HANDLE_BYTES 64
#define SR_EXCHANGE_MAX_RANKS 8
/* handle + byte offset of dev_ptr inside its cudaMalloc allocation (works for pointers into a caching allocator's
 * blocks); open in ANOTHER process -> that process's pointer to the same bytes (peer access enabled lazily). */
int sr_ipc_export(const void* dev_ptr, unsigned char* handle, size_t* offset);
int sr_ipc_open(const unsigned char* handle, size_t offset, void** dev_ptr);
int sr_ipc_close(void* dev_ptr, size_t offset);

typedef struct sr_exchange sr_exchange;
size_t sr_exchange_signal_bytes(void);
/* [lo, hi) in floats of rank's shard of an n-float arena (multiples of 4 except the very end). */
int sr_exchange_shard(size_t n, int rank, int world, size_t* lo, size_t* hi);
/* grads[r], params[r], signals[r]: THIS process's pointers to rank r's arenas / signal pad (r == rank: its own). */
int sr_exchange_create(int rank, int world, size_t n, float* const* grads, float* const* params,
                       void* const* signals, sr_exchange** exchange);
void sr_exchange_destroy(sr_exchange* exchange);
int sr_exchange_set_timeout_ms(sr_exchange* exchange, double ms);
/* sum the ranks' gradients (rank order), Keras-2 Adam with grad_scale (1/world) on the shard, new parameters into
 * every rank's arena; returns when launched, completes when every peer's shard has arrived.  max_blocks 0: default. */
int sr_exchange_adam_step(sr_exchange* exchange, float* m, float* v, int t, float lr, float beta1, float beta2,
                          float eps, float grad_scale, int max_blocks, void* stream);
/* synchronises `stream`; *timed_out = 1 if any wait on a peer expired since creation (results are then invalid). */
int sr_exchange_status(sr_exchange* exchange, void* stream, int* timed_out);
/* Stream-ordered barrier over the ranks' signal pads (each a zeroed buffer of sr_exchange_signal_bytes(), NOT shared
 * with an sr_exchange): when the kernel retires on a rank, everything every rank launched before its own call --
 * stores into peer memory included -- is complete.  Used by the tile-sharded image path (BASELINE config
 * 5): every rank's tail convs store the uint8 pixels their tiles own (img_utils.py:700-722) straight into rank 0's
 * image over NVLink (sr_forward_desc.stitch_u8 = the mapped pointer), then one barrier -- no gather, no merge pass. */
typedef struct sr_peer_barrier sr_peer_barrier;
int sr_peer_barrier_create(int rank, int world, void* const* signals, sr_peer_barrier** barrier);
void sr_peer_barrier_destroy(sr_peer_barrier* barrier);
int sr_peer_barrier_set_timeout_ms(sr_peer_barrier* barrier, double ms);
int sr_peer_barrier_arrive_wait(sr_peer_barrier* barrier, void* stream);
int sr_peer_barrier_status(sr_peer_barrier* barrier, void* stream, int* timed_out);
/* sr_model_apply_gradients with the exchange kernel in place of all-reduce + sr_adam_step: the exchange must have
 * been created over the models' parameter arenas (params[rank] == the pointer given to sr_model_create). */
int sr_model_apply_gradients_exchange(sr_model* model, sr_exchange* exchange, float* m, float* v, int t, float lr,
                                      float beta1, float beta2, float eps, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SR100_H_ */
