"""numpy restatement of the reference's CLI tiling (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows img_utils.extract_patches_Step (img_utils.py:601-676), img_utils.rebuild_from_patches_Step
(img_utils.py:692-724) and the canvas arithmetic + de-processing of
BaseSuperResolutionModel.upscaleStepPatch (models.py:225-256, 336, 351, 391, 412).
Pinned: tests/golden/tiling_ref.npz holds outputs of the reference functions themselves (run verbatim
under stubs by oracle/refgen.py); tests/test_oracle.py checks this file against them bit-exactly.
"""
import numpy as np


def canvas_size(h, w, patch=96, step=64):
    """models.py:225-229 (+patch border) then :248-256 (both dims bumped when either is off-grid)."""
    ch, cw = h + patch, w + patch
    if cw % step != 0 or ch % step != 0:
        cw = int((cw / step) + 1) * step
        ch = int((ch / step) + 1) * step
    return ch, cw


def make_canvas(img_u8, patch=96, step=64):
    h, w = img_u8.shape[:2]
    ch, cw = canvas_size(h, w, patch, step)
    canvas = np.zeros((ch, cw, 3))
    canvas[:h, :w] = img_u8
    return canvas


def extract_patches_step(image, patch_size, step):
    """img_utils.py:622-648: for w in range(i_w-p_w) if w % step == 0: for h in range(i_h-p_h) if h % step == 0."""
    i_h, i_w = image.shape[:2]
    p_h, p_w = patch_size
    if p_h > i_h:
        raise ValueError("Height of the patch should be less than the height of the image.")
    if p_w > i_w:
        raise ValueError("Width of the patch should be less than the width of the image.")
    ws = [w for w in range(i_w - p_w) if w % step == 0]
    hs = [h for h in range(i_h - p_h) if h % step == 0]
    out = np.zeros((len(ws) * len(hs), p_h, p_w, 3))
    n = 0
    for w in ws:
        for h in hs:
            out[n] = image[h:h + p_h, w:w + p_w]
            n += 1
    cnt_h = len(hs) if ws else 0
    return out, (cnt_h, len(ws))


def rebuild_from_patches_step(canvas_shape_hw, patches, patch_size, counts, scale, step, border_crop=8):
    """img_utils.py:702-722: overwrite in w-outer / h-inner order, crop 8 HR px except at index 0."""
    i_h, i_w = canvas_shape_hw
    p_h, p_w = patch_size[0] * scale, patch_size[1] * scale
    cnt_h, cnt_w = counts
    s = step * scale
    out = np.zeros((i_h * scale, i_w * scale, 3))
    n = 0
    for w in range(cnt_w):
        cw = 0 if w == 0 else border_crop
        for h in range(cnt_h):
            ch = 0 if h == 0 else border_crop
            out[h * s + ch:h * s + p_h - ch, w * s + cw:w * s + p_w - cw] = patches[n][ch:p_h - ch, cw:p_w - cw]
            n += 1
    return out


def upscale_step_patch(img_u8, predict, patch=96, step=64, scale=4):
    """models.py:184-415 with `predict` standing in for model.predict: returns (uncropped uint8 canvas,
    cropped uint8 output [0:H*scale, 0:W*scale])."""
    h, w = img_u8.shape[:2]
    canvas = make_canvas(img_u8, patch, step)
    patches, counts = extract_patches_step(canvas, (patch, patch), step)
    x = patches.astype(np.float32) / 255.                      # models.py:336
    y = predict(x).astype(np.float32) * 255.                   # models.py:342, 351
    full = rebuild_from_patches_step(canvas.shape[:2], y, (patch, patch), counts, scale, step)
    full = np.clip(full, 0, 255).astype('uint8')               # models.py:391
    return full, full[0:h * scale, 0:w * scale]                # models.py:412
