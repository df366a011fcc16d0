"""Drop-in mirror of the reference's PSNR.py (PSNR.py:7-109): same function names and return values.
The sum-of-squared-differences reduction runs on the GPU in fp64 (sr_sum_sq_diff_f64); slicing and the
final log10 are host scalars.

Integer inputs keep the reference's dtype arithmetic: `ref_data - target_data` (PSNR.py:14) and `pred - gt`
(PSNR.py:28) are computed in the arrays' own dtype, so uint8 images (what `np.array(PIL image)` gives) wrap
around modulo 256, and PSNRTorch's `imdff ** 2` (int exponent) wraps a second time.  Meaningless as a PSNR, but it
is what the reference returns for those inputs and the goldens of tests/golden/psnr_ref.npz pin it."""
import math

import numpy as np


def _ssd(a, b, int_square=False):
    """sum((a - b)^2).  Floating inputs: fp64 on the device.  Integer inputs: the difference is taken in the
    numpy result dtype (wrap-around) as the reference does; int_square additionally squares in that dtype."""
    from sr100 import ops
    a, b = np.asarray(a), np.asarray(b)
    if np.result_type(a, b).kind in "iub":
        d = a - b
        if int_square:
            return float(np.sum(d ** 2, dtype=np.float64))   # the squares themselves wrapped: no device SSD form
        return ops.sum_sq_diff(d.astype(np.float64), np.zeros(d.shape, dtype=np.float64))
    return ops.sum_sq_diff(a, b)


def psnrVDSR(target, ref, scale):
    """PSNR.py:7-18 (RGB or single-channel arrays, `scale` pixels shaved; scale=0 gives an empty slice
    and NaN in the reference: here a ZeroDivisionError-free NaN as well)."""
    target_data = np.array(target)[scale:-scale, scale:-scale]
    ref_data = np.array(ref)[scale:-scale, scale:-scale]
    if target_data.size == 0:
        return float("nan")
    rmse = math.sqrt(_ssd(ref_data, target_data) / target_data.size)
    return 20 * math.log10(255.0 / rmse)


def PSNRTorch(pred, gt, shave_border=0):
    """PSNR.py:24-32: 20*log10(255/rmse), 100 when identical."""
    height, width = pred.shape[:2]
    pred = pred[shave_border:height - shave_border, shave_border:width - shave_border]
    gt = gt[shave_border:height - shave_border, shave_border:width - shave_border]
    rmse = math.sqrt(_ssd(pred, gt, int_square=True) / pred.size)
    if rmse == 0:
        return 100
    return 20 * math.log10(255.0 / rmse)


def psnrSVLAB(img1, img2):
    """PSNR.py:36-49."""
    img1 = im2double(img1)
    img2 = im2double(img2)
    mse = _ssd(img1, img2) / img1.size
    if mse == 0:
        return 100
    return -10 * math.log10(mse)


def psnrNITRE(pred, gt, shave_border=0):
    """PSNR.py:54-84: inputs with max > 1 are divided by 255; 10*log10(N / sum(diff^2))."""
    height, width = pred.shape[:2]
    pred = pred[shave_border:height - shave_border, shave_border:width - shave_border]
    gt = gt[shave_border:height - shave_border, shave_border:width - shave_border]
    if np.amax(pred) > 1:
        pred = im2double(pred)
    if np.amax(gt) > 1:
        gt = im2double(gt)
    N = pred.size
    sumel = _ssd(pred, gt)
    return 10 * math.log10(N / sumel)


def im2doubleZ(im):
    """PSNR.py:87-91."""
    min_val = np.min(im.ravel())
    max_val = np.max(im.ravel())
    return (im.astype('float') - min_val) / (max_val - min_val)


def im2double(im):
    """PSNR.py:93-98."""
    return np.asarray(im).astype(float) / 255.0


def rgb2y(img):
    """PSNR.py:101-109 is broken in the reference (`y` undefined -> NameError).  This mirror returns the
    intended (H, W, 1) array  y = 16 + 65.481 r + 128.553 g + 24.966 b."""
    r, g, b = img[:, :, 0], img[:, :, 1], img[:, :, 2]
    y = np.zeros(img.shape[:2] + (1,), dtype=np.float64)
    y[:, :, 0] = 16 + (65.481 * r) + (128.553 * g) + (24.966 * b)
    return y
