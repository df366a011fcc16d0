"""Device path of the reference's dataset preparation (img_utils.transform_images, img_utils.py:44-123; SURVEY.md
8f-4).  One input image -> 256 (ground truth, degraded) sub-image pairs, all on the GPU:

    uint8 [h,w,3] -> sr_resize_u8 (Pillow BILINEAR, 256 x 256) -> sr_sharpen3x3_u8 (Pillow SHARPEN)
    -> sr_dataprep_patches: y = bytescale(sample), g = bytescale(gaussian_filter(sample, 0.5))   (one block / sample)
    -> sr_resize_u8 (Pillow BICUBIC, hr -> 16) [-> sr_resize_u8 (BICUBIC, 16 -> hr)]  over all 256 samples at once.

The coefficient tables are computed on the host exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc and
scipy.ndimage._gaussian_kernel1d do; the kernels are integer / float64 restatements of Resample.c, Filter.c,
correlate1d and scipy.misc.bytescale and reproduce the reference function's outputs bit for bit
(tests/test_gpu_dataprep.py against tests/golden/dataprep_ref.npz).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib as L

PRECISION_BITS = 32 - 8 - 2
IMG_SIZE, STRIDE, LR_PATCH = 256, 16, 16       # img_utils.py:23-24, :93 (_image_scale_multiplier = 1)


def _bilinear(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


_FILTERS = {"bilinear": (_bilinear, 1.0), "bicubic": (_bicubic, 2.0), "cubic": (_bicubic, 2.0)}
_coeff_cache = {}


def pil_resize_coeffs(in_size, out_size, interp, device):
    """Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc over the full extent [0, in_size) for the named
    filter -> device tensors (bounds int32 [out,2] = (first tap, taps), kk int32 [out,ksize]), ksize."""
    key = (in_size, out_size, interp, str(device))
    hit = _coeff_cache.get(key)
    if hit is not None:
        return hit
    filt, support = _FILTERS[interp]
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x, v in enumerate(w):
            if ww != 0.0:
                v = v / ww
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    out = (torch.from_numpy(bounds).to(device), torch.from_numpy(kk).to(device), ksize)
    _coeff_cache[key] = out
    return out


def resize_u8(x, out_h, out_w, interp="bilinear"):
    """uint8 device [NB,H,W,3] (or [H,W,3]) -> [NB,out_h,out_w,3]: PIL.Image.resize((out_w, out_h), <interp>)."""
    lib = L.require_device()
    single = x.dim() == 3
    xb = (x.unsqueeze(0) if single else x).contiguous()
    nb, h, w, _ = xb.shape
    bx, kx, ksx = pil_resize_coeffs(w, out_w, interp, xb.device) if out_w != w else (None, None, 0)
    by, ky, ksy = pil_resize_coeffs(h, out_h, interp, xb.device) if out_h != h else (None, None, 0)
    tmp = torch.empty(nb, h, out_w, 3, dtype=torch.uint8, device=xb.device) if (out_w != w and out_h != h) else None
    out = torch.empty(nb, out_h, out_w, 3, dtype=torch.uint8, device=xb.device)
    L.check(lib.sr_resize_u8(L.ptr(xb), nb, h, w, out_h, out_w, L.ptr(bx), L.ptr(kx), ksx, L.ptr(by), L.ptr(ky), ksy,
                             L.ptr(tmp), L.ptr(out), L.stream_ptr()))
    return out[0] if single else out


def sharpen_u8(x):
    """uint8 device [H,W,3] or [NB,H,W,3]: PIL ImageFilter.SHARPEN (scipy.misc.imfilter(img, 'sharpen'))."""
    lib = L.require_device()
    single = x.dim() == 3
    xb = (x.unsqueeze(0) if single else x).contiguous()
    nb, h, w, _ = xb.shape
    out = torch.empty_like(xb)
    L.check(lib.sr_sharpen3x3_u8(L.ptr(xb), nb, h, w, L.ptr(out), L.stream_ptr()))
    return out[0] if single else out


def gaussian_weights(sigma, truncate=4.0):
    """scipy.ndimage._gaussian_kernel1d(sigma, 0, radius), from the centre outwards: (float64 [radius+1], radius)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return phi[radius:].copy(), radius


def subimage_positions(patch, n, img_size=IMG_SIZE, stride=STRIDE):
    """(row, col) of the first n sub-images subimage_generator yields (img_utils.py:134-140): x outer, y inner over
    range(0, img_size - patch, stride); the grid repeats when n exceeds it (it does: 196 positions, 256 samples)."""
    grid = [(x, y) for x in range(0, img_size - patch, stride) for y in range(0, img_size - patch, stride)]
    return np.array([grid[i % len(grid)] for i in range(n)], dtype=np.int32)


def patch_samples(img256, positions, patch, sigma=0.5):
    """-> (y uint8 [n,P,P,3], g uint8 [n,P,P,3]): bytescale(sample) and bytescale(gaussian_filter(sample, sigma))."""
    lib = L.require_device()
    h, w, _ = img256.shape
    n = positions.shape[0]
    pos = torch.from_numpy(np.ascontiguousarray(positions, dtype=np.int32)).to(img256.device)
    wts, radius = gaussian_weights(sigma)
    wd = torch.from_numpy(wts).to(img256.device)
    y = torch.empty(n, patch, patch, 3, dtype=torch.uint8, device=img256.device)
    g = torch.empty_like(y)
    L.check(lib.sr_dataprep_patches(L.ptr(img256.contiguous()), h, w, L.ptr(pos), n, patch, L.ptr(wd), radius,
                                    L.ptr(y), L.ptr(g), L.stream_ptr()))
    return y, g


def transform_image_device(img_u8, scaling_factor=2, true_upscale=False):
    """The body of transform_images' file loop for one decoded RGB image (uint8 device [h,w,3]):
    -> (y uint8 [256,hr,hr,3] ground-truth samples, X uint8 [256,s,s,3] degraded samples), device tensors."""
    hr = 16 * scaling_factor                                   # img_utils.py:77
    n = IMG_SIZE ** 2 // STRIDE ** 2                           # :78
    if hr >= IMG_SIZE or hr > 64:
        raise ValueError("scaling_factor %r: sub-images of %d px are not supported (<= 64)" % (scaling_factor, hr))
    img = resize_u8(img_u8, IMG_SIZE, IMG_SIZE, "bilinear")    # :74
    img = sharpen_u8(img)                                      # :75
    y, g = patch_samples(img, subimage_positions(hr, n), hr)   # :80-103
    x = resize_u8(g, LR_PATCH, LR_PATCH, "bicubic")            # :109
    if not true_upscale:
        x = resize_u8(x, hr, hr, "bicubic")                    # :113
    return y, x
