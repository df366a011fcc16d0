"""CPU restatement of the reference's dataset preparation, img_utils.transform_images (img_utils.py:44-123) and the
library functions it calls.  TEST INFRASTRUCTURE (see oracle/__init__.py): only tests/ may import this.

Per input image (img_utils.py:70-117):
    imread(mode='RGB') -> imresize(img, (256, 256))            scipy.misc default interp='bilinear' = Pillow BILINEAR
    -> scipy.misc.imfilter(img, 'sharpen')                       Pillow ImageFilter.SHARPEN
    -> 256 sub-images of hr = 16 * scaling_factor pixels         subimage_generator (:134-140), stride 16, float64
    per sub-image:
       y  = imsave(ip)                                           toimage(float64) = bytescale (min/max stretch)
       op = gaussian_filter(ip, sigma=0.5)                       scipy.ndimage, ALL THREE axes (channels too), reflect
       op = imresize(op, (16, 16), 'bicubic')                    bytescale, Pillow BICUBIC
       op = imresize(op, (hr, hr), 'bicubic') unless true_upscale
       X  = imsave(op)                                           uint8: no rescale

Third-party arithmetic (none of it under /root/reference; versions unpinned there):
  * Pillow Resample.c 8-bpc resize (oracle/pil_resample.py) with the bilinear / bicubic filters, Filter.c 3x3
    kernel filter (SHARPEN = (-2 ... 32 ... -2) / 16): PINNED bit for bit to the Pillow installed in this image;
  * scipy.ndimage.gaussian_filter (correlate1d, symmetric-kernel summation order): PINNED bit for bit (float64) to
    the scipy installed in this image;
  * scipy.misc.imresize / imfilter / imsave / bytescale: removed from scipy (1.3); restated from the published
    source of scipy 1.2 (scipy/misc/pilutil.py) -- parity unpinned for the glue, pinned for what it calls.
The whole function is additionally pinned to the REFERENCE'S OWN transform_images run under PIL-backed scipy.misc
stubs (oracle/refgen_dataprep.py -> tests/golden/dataprep_ref.npz).
"""
from __future__ import annotations

import numpy as np

from . import pil_resample as pr

IMG_SIZE = 256      # img_utils.py:23 (img_size, _image_scale_multiplier = 1)
STRIDE = 16         # img_utils.py:24


def bilinear_filter(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


FILTERS = {"bilinear": (bilinear_filter, 1.0), "bicubic": (pr.bicubic_filter, 2.0), "cubic": (pr.bicubic_filter, 2.0)}


def resize_u8(img, out_h, out_w, interp="bilinear"):
    """PIL.Image.fromarray(img).resize((out_w, out_h), <interp>) for uint8 [H,W,3]: horizontal pass first."""
    filt, support = FILTERS[interp]
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    cur = img
    if out_w != w:
        cur = pr._pass(cur, *pr.precompute_coeffs(w, out_w, support, filt), axis=1)
    if out_h != h:
        cur = pr._pass(cur, *pr.precompute_coeffs(h, out_h, support, filt), axis=0)
    return cur


def imresize(arr, size, interp="bilinear"):
    """scipy.misc.imresize(arr, (rows, cols), interp): toimage (bytescale unless uint8) -> resize -> uint8 array."""
    return resize_u8(pr.bytescale(arr), int(size[0]), int(size[1]), interp)


def sharpen_u8(img):
    """PIL ImageFilter.SHARPEN on an RGB image (libImaging/Filter.c ImagingFilter3x3, kernel / 16, offset 0):
    interior: clip8(0.5 + 2*c - (sum of the 8 neighbours)/8) with truncation; the one-pixel border is copied.
    Every term is a multiple of 1/8 below 2^11, so the float32 sum is exact in any order."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    out = img.copy()
    if h < 3 or w < 3:
        return out
    a = img.astype(np.int32)
    s9 = np.zeros((h - 2, w - 2, 3), dtype=np.int32)
    for dy in range(3):
        for dx in range(3):
            s9 += a[dy:dy + h - 2, dx:dx + w - 2]
    c = a[1:-1, 1:-1]
    v8 = 16 * c - (s9 - c) + 4                     # 8 * (0.5 + 2c - S8/8)
    val = np.where(v8 <= 0, 0, np.minimum(v8 // 8, 255))
    out[1:-1, 1:-1] = val.astype(np.uint8)
    return out


def gaussian_kernel(sigma, truncate=4.0):
    """scipy.ndimage._gaussian_kernel1d(sigma, 0, radius): exp(-0.5/sigma^2 * x^2) / sum."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum(), radius


def correlate1d_reflect(a, w, radius, axis):
    """NI_Correlate1D, symmetric branch, mode='reflect' (d c b a | a b c d | d c b a):
    tmp = x[0]*w[0]; for j = -radius..-1: tmp += (x[j] + x[-j]) * w[j]   (float64, this order)."""
    a = np.moveaxis(a, axis, -1)
    n = a.shape[-1]
    idx = np.arange(-radius, n + radius)
    period = 2 * n
    idx = np.mod(idx, period)
    idx = np.where(idx >= n, period - 1 - idx, idx)
    ext = a[..., idx]
    c = radius
    out = ext[..., c:c + n] * w[c]
    for j in range(-radius, 0):
        out = out + (ext[..., c + j:c + j + n] + ext[..., c - j:c - j + n]) * w[c + j]
    return np.moveaxis(out, -1, axis)


def gaussian_filter_f64(arr, sigma=0.5):
    """scipy.ndimage.gaussian_filter(arr, sigma) for a float64 array: axes in order 0, 1, 2, ..."""
    w, radius = gaussian_kernel(sigma)
    out = np.asarray(arr, dtype=np.float64)
    for axis in range(out.ndim):
        out = correlate1d_reflect(out, w, radius, axis)
    return out


def subimage_positions(patch, n, img_size=IMG_SIZE, stride=STRIDE):
    """The first n sub-images subimage_generator (img_utils.py:134-140) yields: (x, y) over
    range(0, img_size - patch, stride)^2, x outer, the whole grid repeated as often as needed."""
    grid = [(x, y) for x in range(0, img_size - patch, stride) for y in range(0, img_size - patch, stride)]
    return [grid[i % len(grid)] for i in range(n)]


def transform_image(img_rgb_u8, scaling_factor=2, true_upscale=False):
    """One iteration of the file loop of transform_images: -> (y uint8 [n,hr,hr,3], X uint8 [n,s,s,3])."""
    img = imresize(img_rgb_u8, (IMG_SIZE, IMG_SIZE))                     # img_utils.py:74
    img = sharpen_u8(img)                                                # :75
    hr = 16 * scaling_factor                                             # :77
    n = IMG_SIZE ** 2 // STRIDE ** 2                                     # :78
    lr = 16                                                              # :93
    ys, xs = [], []
    for (x, y) in subimage_positions(hr, n):
        ip = img[x:x + hr, y:y + hr, :].astype(np.float64)               # hr_samples is float64 (:80)
        ys.append(pr.bytescale(ip))                                      # imsave -> toimage -> bytescale (:100)
        op = gaussian_filter_f64(ip, 0.5)                                # :103
        op = imresize(op, (lr, lr), "bicubic")                           # :109
        if not true_upscale:
            op = imresize(op, (hr, hr), "bicubic")                       # :113
        xs.append(op)
    return np.stack(ys), np.stack(xs)
