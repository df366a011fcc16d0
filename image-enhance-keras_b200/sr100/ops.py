"""Tensor-level wrappers of the libsr100 C ABI (device tensors in, device tensors out).

Each function cites the reference operation it replaces; the arithmetic lives in csrc/*.cu.
numpy-facing drop-in signatures are in the top-level mirror modules (img_utils.py, PSNR.py, ...).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _dev():
    L.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def to_device(a, dtype=None):
    """Host array -> device tensor.  Arrays that already live in page-locked memory (e.g. the outputs of
    models.upscale_arrays) are DMA-copied directly; pageable arrays take the ordinary staged copy."""
    a = np.ascontiguousarray(a)
    if not a.flags.writeable:           # e.g. np.asarray(PIL image): torch wants a writable buffer
        a = a.copy()
    t = torch.from_numpy(a)
    if dtype is not None:
        t = t.to(dtype)
    return t.to(_dev(), non_blocking=t.is_pinned())


def to_host_pinned(t):
    """Device tensor -> page-locked host tensor, asynchronously on the current stream (the caller
    synchronises once for all its outputs).  The pinned block comes from torch's caching host allocator and is
    recycled when the returned array is dropped."""
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t, non_blocking=True)
    return out


def patch_count(dim, patch, step):
    """|{x : 0 <= x < dim - patch, x % step == 0}| (img_utils.py:622,629)."""
    return L.load().sr_patch_count(int(dim), int(patch), int(step))


def canvas_size(h, w, patch=96, step=64):
    """Zero-padded canvas of upscaleStepPatch (models.py:225-256)."""
    ch, cw = C.c_int(), C.c_int()
    L.check(L.load().sr_canvas_size(int(h), int(w), int(patch), int(step), C.byref(ch), C.byref(cw)))
    return ch.value, cw.value


def patch_gather_u8(img_u8, canvas_hw, patch, step, divisor=255.0):
    """uint8 [h,w,3] device image -> fp32 patches [N,p,p,3] (extract_patches_Step on the zero-padded
    canvas, then astype(float32)/255., models.py:272,336)."""
    lib = L.require_device()
    h, w, _ = img_u8.shape
    ch, cw = canvas_hw
    ph, pw = patch
    cnt_h, cnt_w = patch_count(ch, ph, step), patch_count(cw, pw, step)
    out = torch.empty(cnt_h * cnt_w, ph, pw, 3, device=img_u8.device, dtype=torch.float32)
    L.check(lib.sr_patch_gather_u8(L.ptr(img_u8), h, w, ch, cw, ph, pw, step, float(divisor), L.ptr(out),
                                   L.stream_ptr()))
    return out, (cnt_h, cnt_w)


def patch_gather_u8_batched(imgs_u8, canvas_hw, patch, step, divisor=255.0, out=None):
    """uint8 [M,h,w,3] device batch of same-shaped images -> fp32 patches [M*N,p,p,3] in one launch (image m's N
    patches follow image m-1's); `out` may be a preallocated slice to fill."""
    lib = L.require_device()
    m, h, w, _ = imgs_u8.shape
    ch, cw = canvas_hw
    ph, pw = patch
    cnt_h, cnt_w = patch_count(ch, ph, step), patch_count(cw, pw, step)
    if out is None:
        out = torch.empty(m * cnt_h * cnt_w, ph, pw, 3, device=imgs_u8.device, dtype=torch.float32)
    L.check(lib.sr_patch_gather_u8_batched(L.ptr(imgs_u8), m, imgs_u8.stride(0), h, w, ch, cw, ph, pw, step,
                                           float(divisor), L.ptr(out), L.stream_ptr()))
    return out, (cnt_h, cnt_w)


def patch_gather_f32(canvas_f32, patch, step):
    """fp32 [H,W,3] device canvas -> fp32 patches (img_utils.extract_patches_Step, img_utils.py:601-676)."""
    lib = L.require_device()
    ch, cw, _ = canvas_f32.shape
    ph, pw = patch
    if ph > ch:
        raise ValueError("Height of the patch should be less than the height of the image.")
    if pw > cw:
        raise ValueError("Width of the patch should be less than the width of the image.")
    cnt_h, cnt_w = patch_count(ch, ph, step), patch_count(cw, pw, step)
    out = torch.empty(cnt_h * cnt_w, ph, pw, 3, device=canvas_f32.device, dtype=torch.float32)
    if out.numel():
        L.check(lib.sr_patch_gather_f32(L.ptr(canvas_f32), ch, cw, ph, pw, step, L.ptr(out), L.stream_ptr()))
    return out, (cnt_h, cnt_w)


def patch_stitch(patches, counts, patch, step, scale, canvas_hw, mul=1.0, want_f32=True, want_u8=False):
    """fp32 patches [N,p*s,p*s,3] -> canvas [H*s,W*s,3] (img_utils.rebuild_from_patches_Step,
    img_utils.py:692-724; optional *255 and clip->uint8 truncation of models.py:351,391)."""
    lib = L.require_device()
    cnt_h, cnt_w = counts
    ph, pw = patch
    ch, cw = canvas_hw
    of = torch.empty(ch * scale, cw * scale, 3, device=patches.device, dtype=torch.float32) if want_f32 else None
    ou = torch.empty(ch * scale, cw * scale, 3, device=patches.device, dtype=torch.uint8) if want_u8 else None
    L.check(lib.sr_patch_stitch(L.ptr(patches), cnt_h, cnt_w, ph, pw, step, scale, ch, cw, float(mul),
                                L.ptr(of), L.ptr(ou), L.stream_ptr()))
    return of, ou


def shard_strip(counts, patch, step, scale, out_w, tile_lo, tile_hi, crop=8):
    """Output columns [x0, x1) that tiles [tile_lo, tile_hi) of the column-major index can own: from the first
    owned column of the first tile column to the last owned column of the last one (ownership passes from tile
    column j-1 to j at scale*step*j + crop, img_utils.py:700-722), clipped to the image."""
    cnt_h, cnt_w = counts
    S, P = step * scale, patch[1] * scale
    j_lo, j_hi = tile_lo // cnt_h, (tile_hi - 1) // cnt_h
    x0 = 0 if j_lo == 0 else S * j_lo + crop
    x1 = S * (j_hi + 1) + crop if j_hi < cnt_w - 1 else S * j_hi + P
    return min(x0, out_w), min(x1, out_w)


def patch_stitch_range(patches, counts, patch, step, scale, canvas_h, tile_lo, tile_hi, x0, strip_w, mul=255.0,
                       out=None):
    """uint8 strip [canvas_h*scale, strip_w, 3] of one shard's owned pixels (sr_patch_stitch_range)."""
    lib = L.require_device()
    cnt_h, cnt_w = counts
    if out is None:
        out = torch.empty(canvas_h * scale, strip_w, 3, device=patches.device, dtype=torch.uint8)
    L.check(lib.sr_patch_stitch_range(L.ptr(patches), cnt_h, cnt_w, patch[0], patch[1], step, scale, canvas_h,
                                      tile_lo, tile_hi, x0, strip_w, float(mul), L.ptr(out), L.stream_ptr()))
    return out


def depth_to_space(x, r, order):
    """fp32 NHWC [N,H,W,C*r*r] -> [N,H*r,W*r,C]; order 0 Subpixel/_phase_shift & depth_to_scale_tf,
    1 depth_to_scale_th, 2 tf.depth_to_space (keras_subpixel.py:64-84, advanced.py:87-129,195-196)."""
    lib = L.require_device()
    n, h, w, c = x.shape
    if c % (r * r):
        raise ValueError("channels must be divisible by r*r")
    out = torch.empty(n, h * r, w * r, c // (r * r), device=x.device, dtype=torch.float32)
    L.check(lib.sr_depth_to_space(L.ptr(x), n, h, w, c // (r * r), r, order, L.ptr(out), L.stream_ptr()))
    return out


def conv2d_direct(x, w_hwio, bias, same=True, relu=False, shuffle_r=0, shuffle_order=0, round_bf16=False):
    """CUDA-core conv, fp32 accumulate (Subpixel(Conv2D) layers outside the 128-channel stack)."""
    lib = L.require_device()
    n, h, wd, cin = x.shape
    k, _, _, cout = w_hwio.shape
    oh, ow = (h, wd) if same else (h - k + 1, wd - k + 1)
    if shuffle_r:
        out = torch.empty(n, oh * shuffle_r, ow * shuffle_r, cout // (shuffle_r * shuffle_r), device=x.device,
                          dtype=torch.float32)
    else:
        out = torch.empty(n, oh, ow, cout, device=x.device, dtype=torch.float32)
    in_is_bf16 = 1 if x.dtype == torch.bfloat16 else 0
    L.check(lib.sr_conv2d_direct(L.ptr(x), in_is_bf16, L.ptr(w_hwio), 1 if round_bf16 else 0, L.ptr(bias), n, h,
                                 wd, cin, cout, k, 1 if same else 0, 1 if relu else 0, shuffle_r, shuffle_order,
                                 L.ptr(out), L.stream_ptr()))
    return out


def conv2d_tc_shuffle(x_bf16, w_hwio, bias, r, order, relu=False):
    """Tensor-core conv (cin = 128, SAME, k in {1,3,5}, cout = r*r*C <= 128) with the depth-to-space shuffle fused
    into the epilogue's store address: bf16 NHWC [N,H,W,128] -> fp32 [N,H*r,W*r,C].  The r*r*C-channel tensor of
    keras_subpixel.Subpixel (keras_subpixel.py:46-61, 109-110) never exists in HBM."""
    lib = L.require_device()
    n, h, w, cin = x_bf16.shape
    k, _, _, cout = w_hwio.shape
    if cin != 128 or x_bf16.dtype != torch.bfloat16:
        raise ValueError("conv2d_tc_shuffle needs a bf16 input with 128 channels")
    packed = torch.empty(lib.sr_packed_weight_bytes(k, cout), dtype=torch.uint8, device=x_bf16.device)
    L.check(lib.sr_pack_conv_weights(L.ptr(w_hwio), k, cout, 0, L.ptr(packed), L.stream_ptr()))
    out = torch.empty(n, h * r, w * r, cout // (r * r), device=x_bf16.device, dtype=torch.float32)
    d = L.ConvDesc()
    d.nsrc = 1
    d.in_[0], d.wpacked[0], d.ksize[0] = x_bf16.data_ptr(), packed.data_ptr(), k
    d.NB, d.H, d.W, d.cin, d.cout = n, h, w, 128, cout
    d.bias = bias.data_ptr() if bias is not None else None
    d.alpha, d.beta, d.relu = 1.0, 0.0, 1 if relu else 0
    d.out_f32 = out.data_ptr()
    d.a_mode, d.nacc, d.pair = 0, 2, 1
    d.shuffle_r, d.shuffle_order = r, order
    plan = C.c_void_p()
    L.check(lib.sr_conv_plan_create(C.byref(d), C.byref(plan)))
    try:
        L.check(lib.sr_conv_plan_run(plan, L.stream_ptr()))
    finally:
        lib.sr_conv_plan_destroy(plan)
    return out


def bilinear4(x, out_dtype=torch.float32):
    """tf.image.resize_bilinear x4, TF1 legacy sampling (models.py:1392-1399)."""
    lib = L.require_device()
    n, h, w, c = x.shape
    out = torch.empty(n, 4 * h, 4 * w, c, device=x.device, dtype=out_dtype)
    ob, of = (out, None) if out_dtype == torch.bfloat16 else (None, out)
    L.check(lib.sr_bilinear4_fwd(L.ptr(x), 1 if x.dtype == torch.bfloat16 else 0, n, h, w, c, L.ptr(ob), L.ptr(of),
                                 L.stream_ptr()))
    return out


def bilinear4_bwd(gout):
    lib = L.require_device()
    n, hh, ww, c = gout.shape
    gin = torch.empty(n, hh // 4, ww // 4, c, device=gout.device, dtype=torch.float32)
    L.check(lib.sr_bilinear4_bwd(L.ptr(gout), n, hh // 4, ww // 4, c, L.ptr(gin), L.stream_ptr()))
    return gin


def rgb2y(img_u8):
    """skimage.color.rgb2ycbcr(im)[:, :, 0] on a uint8 image (scorpath.setimgrgb2ycbcr, scorpath.py:26-31)."""
    lib = L.require_device()
    h, w, _ = img_u8.shape
    y = torch.empty(h, w, device=img_u8.device, dtype=torch.float64)
    L.check(lib.sr_rgb2y_u8(L.ptr(img_u8), h * w, L.ptr(y), L.stream_ptr()))
    return y


def _score_dict(r):
    n = float(r.n_pix)
    # psnrNITRE (PSNR.py:54-84): inputs > 1 are divided by 255, psnr = 10*log10(N / sum(diff^2))
    sum_sq = r.sum_sq_y / (255.0 * 255.0)
    psnr = 10.0 * np.log10(n / sum_sq) if sum_sq > 0 else float("inf")
    ssim_y = r.ssim_y_sum / r.n_win
    ssim_rgb = sum(r.ssim_rgb_sum[i] / r.n_win for i in range(3)) / 3.0
    return dict(psnr_y=float(psnr), ssim_y=float(ssim_y), ssim_rgb=float(ssim_rgb), sum_sq_y=r.sum_sq_y,
                n_pix=r.n_pix, n_win=r.n_win)


def score_pairs(pairs, crop=10):
    """Y-PSNR (psnrNITRE), Y-SSIM and RGB-SSIM of a list of (a, b) pairs of same-shaped uint8 RGB device images
    after a `crop`-pixel border crop (the loop of scorpath.py:92-228): one launch per 32 pairs and ONE read-back
    for all of them.  The pairs may differ in shape.  Returns a list of dicts of python floats."""
    lib = L.require_device()
    pairs = list(pairs)
    if not pairs:
        return []
    items = (L.ScoreItem * len(pairs))()
    for i, (a_u8, b_u8) in enumerate(pairs):
        if a_u8.shape != b_u8.shape:
            raise ValueError("images must have the same shape, got %s and %s" % (tuple(a_u8.shape), tuple(b_u8.shape)))
        items[i].a, items[i].b = L.ptr(a_u8), L.ptr(b_u8)
        items[i].h, items[i].w = int(a_u8.shape[0]), int(a_u8.shape[1])
    size = C.sizeof(L.ScoreResult)
    res = torch.zeros(len(pairs) * size, dtype=torch.uint8, device=pairs[0][0].device)
    L.check(lib.sr_score_batch_u8(C.cast(items, C.c_void_p), len(pairs), crop, L.ptr(res), L.stream_ptr()))
    raw = res.cpu().numpy().tobytes()
    return [_score_dict(L.ScoreResult.from_buffer_copy(raw[i * size:(i + 1) * size])) for i in range(len(pairs))]


def score_pair(a_u8, b_u8, crop=10):
    """Y-PSNR (psnrNITRE), Y-SSIM and RGB-SSIM of two same-shaped uint8 RGB device images after a
    `crop`-pixel border crop (scorpath.py:174-228).  Returns dict of python floats."""
    lib = L.require_device()
    if a_u8.shape != b_u8.shape:
        raise ValueError("images must have the same shape, got %s and %s" % (tuple(a_u8.shape), tuple(b_u8.shape)))
    h, w, _ = a_u8.shape
    res = torch.zeros(C.sizeof(L.ScoreResult), dtype=torch.uint8, device=a_u8.device)
    L.check(lib.sr_score_pair_u8(L.ptr(a_u8), L.ptr(b_u8), h, w, crop, L.ptr(res), L.stream_ptr()))
    return _score_dict(L.ScoreResult.from_buffer_copy(res.cpu().numpy().tobytes()))


def sum_sq_diff(a, b):
    """sum((a-b)^2) in fp64 on the device; a, b numpy arrays of equal shape (any real dtype)."""
    lib = L.require_device()
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        raise ValueError("operands could not be broadcast together with shapes %s %s" % (a.shape, b.shape))
    if a.size == 0:
        return 0.0
    ad, bd = to_device(a), to_device(b)
    out = torch.zeros(1, dtype=torch.float64, device=ad.device)
    L.check(lib.sr_sum_sq_diff_f64(L.ptr(ad), L.ptr(bd), a.size, L.ptr(out), L.stream_ptr()))
    return float(out.item())
