"""CPU restatement of the reference's two alternative tilers with `predict` standing in for model.predict.

TEST INFRASTRUCTURE (see oracle/__init__.py).

  * upscale_patch        BaseSuperResolutionModel.upscalePatch (models.py:419-604): zero-pad to multiples of 4
                         (:461-470), dense uint8 patches (img_utils.make_patchesOrig -> extract_patches_2dv2,
                         img_utils.py:174-180, 561-599), every 4th kept as float64 (extract_patches_2dlocal,
                         img_utils.py:513-556), each shrunk by scipy.misc.imresize(..., 'bicubic') (:487-490; float
                         input => bytescale), /255, predict, *255 (float32), reconstruct_from_patches_2dlocal
                         (img_utils.py:442-511), clip -> uint8, crop (:575-577).
  * upscale_patch_mode   BaseSuperResolutionModel.upscale(mode='patch') (models.py:645-680, 713-791): bicubic x4 of the
                         image, dense uint8 patches (img_utils.make_patches -> sklearn extract_patches_2d), each
                         shrunk x4 (uint8 input => no bytescale), predict, combine_patches = sklearn
                         reconstruct_from_patches_2d (sum in (i, j) order, division by the closed-form overlap count).
The patch loops are the reference's own loops (only the materialisation of ALL dense patches is skipped: patch (i, j)
is sliced from the image when it is needed, which yields the same values).  scipy.misc.imresize -> oracle/pil_resample.py
(bit-exact vs the installed Pillow; bytescale restated from scipy's source).
"""
from __future__ import annotations

from itertools import product

import numpy as np

from . import pil_resample as pr


def pad_to_multiple_of_4(img):
    h, w = img.shape[:2]
    if w % 4 != 0 or h % 4 != 0:                                   # models.py:461-470 (both bumped)
        new_w = int((w / 4) + 1) * 4
        new_h = int((h / 4) + 1) * 4
        new_img = np.zeros((new_h, new_w, 3))
        new_img[0:h, 0:w] = img
        return new_img
    return img


def reconstruct_from_patches_2dlocal(n_hw, patch_hw, patchcnn, image_size, step):
    """img_utils.py:442-511 (the dense `patches` argument only provides the (i, j) enumeration and the patch shape)."""
    i_h, i_w = image_size[:2]
    p_h, p_w = patch_hw
    img = np.zeros(image_size)
    imgmap = np.zeros(image_size)
    n_h, n_w = n_hw
    cnt = 0
    pad = 4
    for i, j in product(range(n_h), range(n_w)):
        if i % step == 0 and j % step == 0:
            if i > 0 and j > 0 and i < n_h - 1 and j < n_w - 1:
                pa = patchcnn[cnt]
                img[i + pad:i + p_h - pad, j + pad:j + p_w - pad] += pa[pad:p_h - pad, pad:p_w - pad]
                imgmap[i + pad:i + p_h - pad, j + pad:j + p_w - pad] += 1
            else:
                img[i:i + p_h, j:j + p_w] += patchcnn[cnt]
                imgmap[i:i + p_h, j:j + p_w] += 1
            cnt += 1
    with np.errstate(invalid="ignore", divide="ignore"):
        return img / imgmap


def upscale_patch(img_u8, predict, patch_size=32, scalemulti=4):
    orig_h, orig_w = img_u8.shape[:2]
    true_img = pad_to_multiple_of_4(img_u8)
    H, W = true_img.shape[:2]
    p = patch_size
    if p > H:                                                       # extract_patches_2dv2, img_utils.py:566-572
        raise ValueError("Height of the patch should be less than the height of the image.")
    if p > W:
        raise ValueError("Width of the patch should be less than the width of the image.")
    n_h, n_w = H - p + 1, W - p + 1
    u8 = true_img.astype('uint8')                                   # extract_patches_2dv2: image.astype('uint8')
    sel = [(i, j) for i, j in product(range(n_h), range(n_w)) if i % 4 == 0 and j % 4 == 0]
    q = int(p / scalemulti)
    small = np.zeros((len(sel), q, q, 3), dtype=np.float32)
    for n, (i, j) in enumerate(sel):
        patchel = u8[i:i + p, j:j + p].astype(np.float64)          # new_patch is a float64 array (img_utils.py:544)
        small[n] = pr.imresize_bicubic(patchel, (q, q))
    result = predict(small.astype(np.float32) / 255.).astype(np.float32) * 255.
    out = reconstruct_from_patches_2dlocal((n_h, n_w), (p, p), result, (H, W, 3), 4)
    out = np.clip(out, 0, 255).astype('uint8')
    return out[0:orig_h, 0:orig_w]


def upscale_patch_mode(img_u8, predict, patch_size=32):
    rows, cols = img_u8.shape[:2]
    big = pr.resize_bicubic_u8(img_u8, rows * 4, cols * 4)          # imresize(true_img, (init_width*4, init_height*4))
    p = patch_size
    H, W = big.shape[:2]
    n_h, n_w = H - p + 1, W - p + 1
    q = int(p / 4)
    small = np.zeros((n_h * n_w, q, q, 3), dtype=np.float32)
    for n, (i, j) in enumerate(product(range(n_h), range(n_w))):
        small[n] = pr.imresize_bicubic(big[i:i + p, j:j + p], (q, q))
    result = predict(small.astype(np.float32) / 255.).astype(np.float32) * 255.
    img = np.zeros((H, W, 3))                                       # sklearn reconstruct_from_patches_2d
    for pch, (i, j) in zip(result, product(range(n_h), range(n_w))):
        img[i:i + p, j:j + p] += pch
    for i in range(H):
        for j in range(W):
            img[i, j] /= float(min(i + 1, p, H - i) * min(j + 1, p, W - j))
    return big, np.clip(img, 0, 255).astype('uint8')
