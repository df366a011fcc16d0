"""Drop-in mirror of the function API of the reference's imgpatch.py (vendored scikit-learn patch helpers with
the author's step-subsampled variants, imgpatch.py:24-338).  No import-time side effects: the reference's
module-level script (imgpatch.py:19-20, 341-358: os.listdir of hard-coded /home/www paths) is dropped.

These helpers are NOT on the CLI hot path (main_dirpath.py uses img_utils.extract_patches_Step /
rebuild_from_patches_Step, which run on the GPU).  Everything here is index arithmetic on the host: dense patch k of
an (i_h, i_w) image sits at (k // n_w, k % n_w) with n_w = i_w - p_w + 1, the step-subsampled variants keep the
positions whose row AND column are multiples of `step`, in that order.  Outputs are pinned to the reference's own
functions (tests/golden/imgpatch_ref.npz) and to the known answer in its docstring (imgpatch.py:193-213).
"""
import numbers

import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

_TOO_TALL = "Height of the patch should be less than the height of the image."
_TOO_WIDE = "Width of the patch should be less than the width of the image."


def _step_positions(image_hw, patch_hw, step, limit):
    """(dense index k, row i, column j) of the dense patches with i % step == 0 and j % step == 0, in dense order,
    cut off at `limit` dense patches (the reference zips the patch array with the dense enumeration)."""
    n_h, n_w = image_hw[0] - patch_hw[0] + 1, image_hw[1] - patch_hw[1] + 1
    if n_h <= 0 or n_w <= 0:
        return []
    rows, cols = np.arange(0, n_h, step), np.arange(0, n_w, step)
    k = (rows[:, None] * n_w + cols[None, :]).ravel()
    k = k[k < limit]
    return [(int(v), int(v // n_w), int(v % n_w)) for v in k]


def reconstruct_from_patches_2d(patches, image_size, step=16):
    """imgpatch.py:24-62.  Despite the sklearn docstring the reference OVERWRITES (the averaging loop sits after
    `return img`): the dense patch at (i, j) lands on the image iff both are multiples of `step`, later ones on top."""
    p_h, p_w = patches.shape[1:3]
    img = np.zeros(image_size)
    for k, i, j in _step_positions(image_size[:2], (p_h, p_w), step, len(patches)):
        img[i:i + p_h, j:j + p_w] = patches[k]
    return img


def _compute_n_patches(i_h, i_w, p_h, p_w, max_patches=None):
    """imgpatch.py:77-109: all dense patches, an integer cap, or a fraction of them."""
    total = (i_h - p_h + 1) * (i_w - p_w + 1)
    if not max_patches:
        return total
    if isinstance(max_patches, numbers.Integral) and max_patches < total:
        return max_patches
    if isinstance(max_patches, numbers.Real) and 0 < max_patches < 1:
        return int(max_patches * total)
    raise ValueError("Invalid value for max_patches: %r" % max_patches)


def extract_patches(arr, patch_shape=8, extraction_step=1):
    """imgpatch.py:113-161: the 2n-dimensional strided VIEW (no copy) of all patches, every `extraction_step`-th one
    per axis: result[i0, i1, ..., :, :, ...] = arr[i0*s0 : i0*s0 + p0, ...]."""
    nd = arr.ndim
    patch_shape = (patch_shape,) * nd if isinstance(patch_shape, numbers.Number) else tuple(patch_shape)
    steps = (extraction_step,) * nd if isinstance(extraction_step, numbers.Number) else tuple(extraction_step)
    windows = sliding_window_view(arr, patch_shape)            # index axes first, patch axes last
    return windows[tuple(slice(None, None, s) for s in steps)]


def extract_patches_2d(image, patch_size, max_patches=None, random_state=None):
    """imgpatch.py:164-248 (known answer in its docstring, imgpatch.py:193-213)."""
    i_h, i_w = image.shape[:2]
    p_h, p_w = patch_size
    if p_h > i_h:
        raise ValueError(_TOO_TALL)
    if p_w > i_w:
        raise ValueError(_TOO_WIDE)
    image = np.asarray(image).reshape((i_h, i_w, -1))
    n_colors = image.shape[-1]
    dense = extract_patches(image, patch_shape=(p_h, p_w, n_colors), extraction_step=1)
    n_patches = _compute_n_patches(i_h, i_w, p_h, p_w, max_patches)
    if max_patches:
        rng = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)
        rows = rng.randint(i_h - p_h + 1, size=n_patches)       # rows first, then columns: the reference's draw order
        cols = rng.randint(i_w - p_w + 1, size=n_patches)
        dense = dense[rows, cols, 0]
    out = dense.reshape(-1, p_h, p_w, n_colors)
    return out.reshape((n_patches, p_h, p_w)) if n_colors == 1 else out


def reconstruct_from_patches_2dlocal(patches, patchcnn, image_size, step=16):
    """imgpatch.py:250-293: positions come from the DENSE enumeration of `patches`; the pixels written are the
    step-selected `patchcnn[0], patchcnn[1], ...` in the same order (overwrite)."""
    p_h, p_w = patches.shape[1:3]
    img = np.zeros(image_size)
    for n, (_, i, j) in enumerate(_step_positions(image_size[:2], (p_h, p_w), step, len(patches))):
        img[i:i + p_h, j:j + p_w] = patchcnn[n]
    return img


def extract_patches_2dlocal(image, patches, patch_size, step=None):
    """imgpatch.py:295-337: the dense patches whose (i, j) are both multiples of `step`, as float64 [n, p_h, p_w, 3]."""
    i_h, i_w = image.shape[:2]
    if patch_size[0] > i_h:
        raise ValueError(_TOO_TALL)
    if patch_size[1] > i_w:
        raise ValueError(_TOO_WIDE)
    p_h, p_w = patches.shape[1:3]
    keep = [k for k, _, _ in _step_positions((i_h, i_w), (p_h, p_w), step, len(patches))]
    out = np.zeros((len(keep), p_h, p_w, 3))
    if keep:
        out[:] = np.asarray(patches)[keep]
    return out
