"""Device path of the reference's two alternative tilers:

  * BaseSuperResolutionModel.upscalePatch (models.py:419-604), "enhance at the same size": every 4th p x p patch of
    the image -> scipy.misc.imresize(patch, (p/4, p/4), 'bicubic') (bytescale + Pillow bicubic) -> /255 -> network
    (x4, back to p x p) -> x255 -> img_utils.reconstruct_from_patches_2dlocal averaging (4-px interior crop, count
    map) -> clip -> uint8;
  * BaseSuperResolutionModel.upscale(mode='patch') (models.py:645-680, 758-790), "x4": the image is first enlarged
    x4 with PIL bicubic, then EVERY dense p x p patch of it goes through the same shrink -> network path and the
    outputs are averaged by sklearn's reconstruct_from_patches_2d (img_utils.combine_patches).

One grid row of patches at a time: `sr_patch_down4_u8` -> `engine.forward_device` -> `sr_patch_average_accumulate`
(float64 sums in the reference's patch order), then `sr_patch_average_finalize`.  The reference materialises all
dense patches first ((H-p+1)(W-p+1) x p x p x 3 bytes, img_utils.make_patchesOrig / make_patches); here a row band
of patches exists at any time.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib as L

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_coeffs(in_size, out_size):
    """Pillow's Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (support 2) over the full
    extent [0, in_size): (bounds int32 [out,2] = (first tap, taps), kk int32 [out, ksize], 22 fractional bits)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x, v in enumerate(w):
            if ww != 0.0:
                v = v / ww
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def patch_down4(img_u8, p, step, stretch, n0=None, n1=None, coeffs=None):
    """uint8 device image [H,W,3] -> float32 [n1-n0, p/4, p/4, 3] in [0,1]: grid patches n0..n1-1 shrunk x4."""
    lib = L.require_device()
    H, W, _ = img_u8.shape
    cnt_h, cnt_w = (H - p) // step + 1, (W - p) // step + 1
    n0 = 0 if n0 is None else n0
    n1 = cnt_h * cnt_w if n1 is None else n1
    if coeffs is None:
        b, k = pil_bicubic_coeffs(p, p // 4)
        coeffs = (torch.from_numpy(b).to(img_u8.device), torch.from_numpy(k).to(img_u8.device))
    bd, kd = coeffs
    q = p // 4
    out = torch.empty(n1 - n0, q, q, 3, device=img_u8.device, dtype=torch.float32)
    L.check(lib.sr_patch_down4_u8(L.ptr(img_u8), H, W, p, step, cnt_h, cnt_w, n0, n1, 1 if stretch else 0,
                                  L.ptr(bd), L.ptr(kd), int(kd.shape[1]), 255.0, L.ptr(out), L.stream_ptr()))
    return out


def patch_average(chunks, P, step, pad, cnt_h, cnt_w, out_hw, mul=255.0, want_f64=False, sklearn_count=False,
                  edges=None):
    """chunks: iterable of (a0, a1, float32 [(a1-a0)*cnt_w, P, P, 3]) in increasing a0 -> (uint8 image, float64 image).
    sklearn_count: divide by reconstruct_from_patches_2d's closed-form overlap count instead of the count map.
    edges: grid indices (a, b) of the patches on the last dense position (the reference's border test); default: the
    last row / column of the grid."""
    lib = L.require_device()
    out_h, out_w = out_hw
    dev = torch.device("cuda", torch.cuda.current_device())
    acc = torch.zeros(out_h, out_w, 3, device=dev, dtype=torch.float64)
    cnt = torch.zeros(out_h, out_w, device=dev, dtype=torch.int32)
    st = L.stream_ptr()
    ea, eb = (cnt_h - 1, cnt_w - 1) if edges is None else edges
    for a0, a1, patches in chunks:
        L.check(lib.sr_patch_average_accumulate(L.ptr(patches), P, step, pad, cnt_h, cnt_w, a0, a1, ea, eb, float(mul),
                                                out_h, out_w, L.ptr(acc), L.ptr(cnt), st))
    u8 = torch.empty(out_h, out_w, 3, device=dev, dtype=torch.uint8)
    f64 = torch.empty(out_h, out_w, 3, device=dev, dtype=torch.float64) if want_f64 else None
    L.check(lib.sr_patch_average_finalize(L.ptr(acc), L.ptr(cnt), out_h, out_w, P if sklearn_count else 0, L.ptr(f64),
                                          L.ptr(u8), st))
    return u8, f64


def enhance_image_device(engine, img_u8, p, step, stretch, pad, max_patches=8192):
    """img_u8: uint8 device image [H,W,3] with (H-p) % step == (W-p) % step == 0 -> uint8 [H,W,3]: every grid patch
    shrunk x4, run through `engine` (a x4 model), averaged back."""
    H, W, _ = img_u8.shape
    if p > H:
        raise ValueError("Height of the patch should be less than the height of the image.")
    if p > W:
        raise ValueError("Width of the patch should be less than the width of the image.")
    if getattr(engine, "scale", 4) != 4:
        raise ValueError("the patch tilers shrink every patch x4: they need a x4 model")
    cnt_h, cnt_w = (H - p) // step + 1, (W - p) // step + 1
    b, k = pil_bicubic_coeffs(p, p // 4)
    coeffs = (torch.from_numpy(b).to(img_u8.device), torch.from_numpy(k).to(img_u8.device))
    rows = max(1, max_patches // cnt_w)

    def chunks():
        for a0 in range(0, cnt_h, rows):
            a1 = min(cnt_h, a0 + rows)
            x = patch_down4(img_u8, p, step, stretch, a0 * cnt_w, a1 * cnt_w, coeffs)
            yield a0, a1, engine.forward_device(x)

    u8, _ = patch_average(chunks(), p, step, pad, cnt_h, cnt_w, (H, W), sklearn_count=(step == 1 and pad == 0))
    return u8
