"""numpy/scipy restatement of the reference's scoring path (TEST INFRASTRUCTURE, see oracle/__init__.py).

PSNR functions follow PSNR.py:7-98 and are pinned by tests/golden/psnr_ref.npz (the reference's PSNR.py run
verbatim).  rgb2ycbcr / compare_ssim restate scikit-image (skimage.color.rgb2ycbcr, and
skimage.measure.compare_ssim as called at scorpath.py:226,228: win_size 7, uniform window, sample
covariance, K1=0.01, K2=0.03, data_range=255, mean over the map cropped by 3 px; multichannel = mean of
the per-channel results); skimage is not installable here, so that part is PARITY UNPINNED.
"""
import math

import numpy as np
from scipy.ndimage import uniform_filter


def crop_border(img, b):
    """scorpath.py:67-70."""
    return img[b:img.shape[0] - b, b:img.shape[1] - b]


def rgb2ycbcr_y(img_u8):
    """skimage.color.rgb2ycbcr(img)[..., 0]: img_as_float (x/255) @ [65.481, 128.553, 24.966] + 16."""
    arr = img_u8.astype(np.float64) / 255.0
    return arr[..., 0] * 65.481 + arr[..., 1] * 128.553 + arr[..., 2] * 24.966 + 16.0


def im2double(im):
    return im.astype(np.float64) / 255.0


def psnr_nitre(pred, gt, shave=0):
    """PSNR.py:54-84."""
    h, w = pred.shape[:2]
    pred = pred[shave:h - shave, shave:w - shave]
    gt = gt[shave:h - shave, shave:w - shave]
    if np.amax(pred) > 1:
        pred = im2double(pred)
    if np.amax(gt) > 1:
        gt = im2double(gt)
    d = pred - gt
    return 10 * math.log10(d.size / np.sum(d ** 2))


def psnr_torch(pred, gt, shave=0):
    """PSNR.py:24-32."""
    h, w = pred.shape[:2]
    pred = pred[shave:h - shave, shave:w - shave]
    gt = gt[shave:h - shave, shave:w - shave]
    rmse = math.sqrt(np.mean((pred.astype(np.float64) - gt.astype(np.float64)) ** 2))
    return 100 if rmse == 0 else 20 * math.log10(255.0 / rmse)


def psnr_vdsr(target, ref, scale):
    """PSNR.py:7-18."""
    t = np.array(target)[scale:-scale, scale:-scale].astype(np.float64)
    r = np.array(ref)[scale:-scale, scale:-scale].astype(np.float64)
    return 20 * math.log10(255.0 / math.sqrt(np.mean((r - t).flatten('C') ** 2.)))


def psnr_svlab(a, b):
    """PSNR.py:36-49."""
    mse = np.mean((im2double(a) - im2double(b)) ** 2)
    return 100 if mse == 0 else -10 * math.log10(mse)


def ssim_single(x, y, data_range=255.0, win=7):
    x = x.astype(np.float64)
    y = y.astype(np.float64)
    npx = win * win
    cov_norm = npx / (npx - 1.0)                      # use_sample_covariance=True
    ux, uy = uniform_filter(x, win), uniform_filter(y, win)
    uxx, uyy, uxy = uniform_filter(x * x, win), uniform_filter(y * y, win), uniform_filter(x * y, win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return float(s[pad:s.shape[0] - pad, pad:s.shape[1] - pad].mean())


def ssim(x, y, data_range=255.0, multichannel=False):
    if multichannel:
        return float(np.mean([ssim_single(x[..., c], y[..., c], data_range) for c in range(x.shape[-1])]))
    return ssim_single(x, y, data_range)


def score_pair(im1_u8, im2_u8, crop=10):
    """scorpath.py:174-228 for one (GT, SR) pair: (psnrNITRE on Y, SSIM on RGB, SSIM on Y)."""
    a, b = crop_border(im1_u8, crop), crop_border(im2_u8, crop)
    ya, yb = rgb2ycbcr_y(a), rgb2ycbcr_y(b)
    return psnr_nitre(yb, ya, 0), ssim(a, b, 255.0, multichannel=True), ssim(ya, yb, 255.0)
