"""numpy restatement of Pillow's 8-bit-per-channel bicubic resize (libImaging/Resample.c: precompute_coeffs,
normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc / Vertical_8bpc) and of scipy.misc.imresize / bytescale
(scipy < 1.3, scipy/misc/pilutil.py) which the reference calls as `imresize(..., interp='bicubic')`
(models.py:490, 655, 672; img_utils.py:107-111).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned: `resize_bicubic_u8` is checked bit for bit against the
Pillow installed in this image (tests/test_alt_tilers.py); scipy.misc itself is absent (removed in scipy 1.3), so
`imresize`'s bytescale step is restated from its published source - parity unpinned for that step.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size, support=2.0, filt=bicubic_filter):
    """-> (bounds [out,2] (xmin, count), kk int32 [out, ksize]) for resizing the full extent [0, in_size)."""
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
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """one 8bpc pass along `axis` (0 = vertical, 1 = horizontal) of a uint8 [H,W,C] image"""
    src = img.astype(np.int64)
    out_n = bounds.shape[0]
    shape = list(img.shape)
    shape[axis] = out_n
    out = np.zeros(shape, dtype=np.uint8)
    for xx in range(out_n):
        xmin, cnt = bounds[xx]
        acc = np.full(shape[:axis] + shape[axis + 1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(cnt):
            acc = acc + np.take(src, xmin + x, axis=axis) * int(kk[xx, x])
        v = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
        if axis == 0:
            out[xx] = v
        else:
            out[:, xx] = v
    return out


def resize_bicubic_u8(img, out_h, out_w):
    """PIL.Image.fromarray(img).resize((out_w, out_h), BICUBIC): horizontal pass first, uint8 intermediate."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    cur = img
    if out_w != w:
        cur = _pass(cur, *precompute_coeffs(w, out_w), axis=1)
    if out_h != h:
        cur = _pass(cur, *precompute_coeffs(h, out_h), axis=0)
    return cur


def bytescale(data, cmin=None, cmax=None, high=255, low=0):
    """scipy.misc.bytescale: uint8 passes through; anything else is stretched from [min, max] to [0, 255]."""
    data = np.asarray(data)
    if data.dtype == np.uint8:
        return data
    if cmin is None:
        cmin = data.min()
    if cmax is None:
        cmax = data.max()
    cscale = cmax - cmin
    if cscale < 0:
        raise ValueError("`cmax` should be larger than `cmin`.")
    elif cscale == 0:
        cscale = 1
    scale = float(high - low) / cscale
    bytedata = (data - cmin) * scale + low
    return (bytedata.clip(low, high) + 0.5).astype(np.uint8)


def imresize_bicubic(arr, size):
    """scipy.misc.imresize(arr, size=(rows, cols), interp='bicubic') for [H,W,3] arrays: toimage (bytescale) ->
    PIL resize((cols, rows), BICUBIC) -> uint8 array."""
    return resize_bicubic_u8(bytescale(arr), int(size[0]), int(size[1]))
