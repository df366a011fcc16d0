"""Drop-in mirror of the reference's scorpath.py: colour helpers (scorpath.py:26-74) and the directory
scoring loop (scorpath.py:76-258).  The per-pair work (border crop, RGB->Y, Y-PSNR, Y-SSIM, RGB-SSIM) is
one fused libsr100 kernel (csrc/score.cu); cv2-based helpers call cv2 exactly as the reference does."""
from __future__ import print_function, division

import os
import sys

import numpy as np

from PSNR import psnrVDSR, PSNRTorch, psnrNITRE, psnrSVLAB, im2double  # noqa: F401  (re-exported like the reference)

DEFAULT_PATH_DIR = "/home/www/imgsuper/val_images/set5nitre/"  # scorpath.py:93


def setimgrgb2ycbcr(im):
    """scorpath.py:26-31: skimage.color.rgb2ycbcr(im)[:, :, 0] (float64, [16, 235]) on the GPU."""
    import torch
    from sr100 import ops
    im = np.asarray(im)
    if im.dtype != np.uint8:
        # skimage img_as_float leaves float input unscaled; the studio-range formula still applies
        imf = im.astype(np.float64)
        return 16.0 + (imf[..., 0] * 65.481 + imf[..., 1] * 128.553 + imf[..., 2] * 24.966)
    return ops.rgb2y(ops.to_device(im, torch.uint8)).cpu().numpy()


def rgb2ycbcrLocal(im):
    """scorpath.py:34-38 (JPEG full-range matrix, np.uint8 truncation)."""
    xform = np.array([[.299, .587, .114], [-.1687, -.3313, .5], [.5, -.4187, -.0813]])
    ycbcr = im.dot(xform.T)
    ycbcr[:, :, [1, 2]] += 128
    return np.uint8(ycbcr)


def rgb2ycbcrTORCH(im):
    """scorpath.py:40-43."""
    im = im2double(im)
    y = 16 + (65.481 * im[:, :, 0]) + (128.553 * im[:, :, 1]) + (24.966 * im[:, :, 2])
    return y.astype(np.float32)


def rgb2ycbcrCV(im_rgb):
    """scorpath.py:48-54."""
    import cv2
    im_rgb = im_rgb.astype(np.float32)
    im_ycrcb = cv2.cvtColor(im_rgb, cv2.COLOR_RGB2YCR_CB)
    im_ycbcr = im_ycrcb[:, :, (0, 2, 1)].astype(np.float32)
    im_ycbcr[:, :, 0] = (im_ycbcr[:, :, 0] * (235 - 16) + 16) / 255.0
    im_ycbcr[:, :, 1:] = (im_ycbcr[:, :, 1:] * (240 - 16) + 16) / 255.0
    return im_ycbcr


def ycbcr2rgb(im_ycbcr):
    """scorpath.py:56-62."""
    import cv2
    im_ycbcr = im_ycbcr.astype(np.float32)
    im_ycbcr[:, :, 0] = (im_ycbcr[:, :, 0] * 255.0 - 16) / (235 - 16)
    im_ycbcr[:, :, 1:] = (im_ycbcr[:, :, 1:] * 255.0 - 16) / (240 - 16)
    im_ycrcb = im_ycbcr[:, :, (0, 2, 1)].astype(np.float32)
    return cv2.cvtColor(im_ycrcb, cv2.COLOR_YCR_CB2RGB)


def crop_border(imgage, bordr):
    """scorpath.py:67-70."""
    init_width, init_height = imgage.shape[0], imgage.shape[1]
    return imgage[bordr: init_width - bordr, bordr: init_height - bordr]


def im2double1(im):
    """scorpath.py:71-74."""
    return im.astype(float) / 255.0


def _imread_rgb(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"))


def score_pair(im1, im2, cropval=10):
    """One iteration of scorpath.py:174-228 on two same-shaped uint8 RGB images:
    returns (psnrNITRE on Y, SSIM on RGB, SSIM on Y)."""
    import torch
    from sr100 import ops
    r = ops.score_pair(ops.to_device(np.asarray(im1), torch.uint8), ops.to_device(np.asarray(im2), torch.uint8),
                       crop=cropval)
    return r["psnr_y"], r["ssim_rgb"], r["ssim_y"]


def score_pairs(pairs, cropval=10):
    """score_pair for a list of (im1, im2) pairs (shapes may differ from pair to pair): one launch and one
    read-back for the whole list.  Returns a list of (psnrNITRE on Y, SSIM on RGB, SSIM on Y)."""
    import torch
    from sr100 import ops
    dev = [(ops.to_device(np.asarray(a), torch.uint8), ops.to_device(np.asarray(b), torch.uint8)) for a, b in pairs]
    return [(r["psnr_y"], r["ssim_rgb"], r["ssim_y"]) for r in ops.score_pairs(dev, crop=cropval)]


def main(path_dir=DEFAULT_PATH_DIR, suffix='scaled', scale_factor=1):
    """scorpath.py:76-258: pair every file without `suffix` in its name with
    '<stem>_<suffix>(<scale_factor>x)<ext>', score, print means.  The pairs of the directory are scored in one
    batched launch; the per-file prints keep the reference's order and wording."""
    scorlist, scorssimy, scorski = [], [], []
    names, pairs = [], []
    for file in os.listdir(path_dir):
        pathfile = path_dir + file
        path = os.path.splitext(pathfile)
        if suffix not in pathfile:
            fileOrig = path[0] + path[1]
            filenameNitre = path[0] + "_" + suffix + "(%dx)" % (scale_factor) + path[1]
            names.append((fileOrig, filenameNitre))
            pairs.append((_imread_rgb(fileOrig), _imread_rgb(filenameNitre)))
    for (fileOrig, filenameNitre), (scor, ski, ski_y) in zip(names, score_pairs(pairs, 10)):
        print(fileOrig)
        print(filenameNitre)
        print("SCORs psnr_ski")
        print(scor)
        scorlist.append(scor)
        scorski.append(ski)
        scorssimy.append(ski_y)
        print("SCORs SSIM Y")
        print(ski_y)
    meanPNSR = sum(scorlist) / float(len(scorlist))
    meanSKI = sum(scorski) / float(len(scorski))
    meanY = sum(scorssimy) / float(len(scorssimy))
    print("-" * 79)
    print("SCOR MEAN psnr")
    print(meanPNSR)
    print("SCOR MEAN SSIM SKI")
    print(meanSKI)
    print("SCOR MEAN SSIM SKI y")
    print(meanY)
    return meanPNSR, meanSKI, meanY


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else DEFAULT_PATH_DIR)
