"""Drop-in mirror of the function API of the reference's imgpatch.py (vendored scikit-learn patch helpers with
the author's step-subsampled variants, imgpatch.py:24-338).  No import-time side effects: the reference's
module-level script (imgpatch.py:19-20, 341-358: os.listdir of hard-coded /home/www paths) is dropped.

These helpers are NOT on the CLI hot path (main_dirpath.py uses img_utils.extract_patches_Step /
rebuild_from_patches_Step, which run on the GPU).  extract_patches is an O(1) strided VIEW exactly as in the
reference; the reconstruct_* functions are index glue kept on the host in vectorised numpy form.
"""
import numbers
from itertools import product

import numpy as np
from numpy.lib.stride_tricks import as_strided


def reconstruct_from_patches_2d(patches, image_size, step=16):
    """imgpatch.py:24-62: despite the sklearn docstring, the reference OVERWRITES (no averaging; the division
    loop after `return img` is dead code): dense patch (i, j) is written iff i % step == 0 and j % step == 0."""
    i_h, i_w = image_size[:2]
    p_h, p_w = patches.shape[1:3]
    img = np.zeros(image_size)
    n_h = i_h - p_h + 1
    n_w = i_w - p_w + 1
    for p, (i, j) in zip(patches, product(range(n_h), range(n_w))):
        if i % step == 0 and j % step == 0:
            img[i:i + p_h, j:j + p_w] = p
    return img


def _compute_n_patches(i_h, i_w, p_h, p_w, max_patches=None):
    """imgpatch.py:77-109."""
    n_h = i_h - p_h + 1
    n_w = i_w - p_w + 1
    all_patches = n_h * n_w
    if max_patches:
        if isinstance(max_patches, numbers.Integral) and max_patches < all_patches:
            return max_patches
        elif isinstance(max_patches, numbers.Real) and 0 < max_patches < 1:
            return int(max_patches * all_patches)
        else:
            raise ValueError("Invalid value for max_patches: %r" % max_patches)
    else:
        return all_patches


def extract_patches(arr, patch_shape=8, extraction_step=1):
    """imgpatch.py:113-161: 2n-dimensional strided view (no copy)."""
    arr_ndim = arr.ndim
    if isinstance(patch_shape, numbers.Number):
        patch_shape = tuple([patch_shape] * arr_ndim)
    if isinstance(extraction_step, numbers.Number):
        extraction_step = tuple([extraction_step] * arr_ndim)
    patch_strides = arr.strides
    slices = tuple(slice(None, None, st) for st in extraction_step)
    indexing_strides = arr[slices].strides
    patch_indices_shape = ((np.array(arr.shape) - np.array(patch_shape)) // np.array(extraction_step)) + 1
    shape = tuple(list(patch_indices_shape) + list(patch_shape))
    strides = tuple(list(indexing_strides) + list(patch_strides))
    return as_strided(arr, shape=shape, strides=strides)


def extract_patches_2d(image, patch_size, max_patches=None, random_state=None):
    """imgpatch.py:164-248 (known answer in its docstring, imgpatch.py:193-213)."""
    i_h, i_w = image.shape[:2]
    p_h, p_w = patch_size
    if p_h > i_h:
        raise ValueError("Height of the patch should be less than the height"
                         " of the image.")
    if p_w > i_w:
        raise ValueError("Width of the patch should be less than the width"
                         " of the image.")
    image = np.asarray(image)
    image = image.reshape((i_h, i_w, -1))
    n_colors = image.shape[-1]
    extracted_patches = extract_patches(image, patch_shape=(p_h, p_w, n_colors), extraction_step=1)
    n_patches = _compute_n_patches(i_h, i_w, p_h, p_w, max_patches)
    if max_patches:
        rng = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)
        i_s = rng.randint(i_h - p_h + 1, size=n_patches)
        j_s = rng.randint(i_w - p_w + 1, size=n_patches)
        patches = extracted_patches[i_s, j_s, 0]
    else:
        patches = extracted_patches
    patches = patches.reshape(-1, p_h, p_w, n_colors)
    if patches.shape[-1] == 1:
        return patches.reshape((n_patches, p_h, p_w))
    return patches


def reconstruct_from_patches_2dlocal(patches, patchcnn, image_size, step=16):
    """imgpatch.py:250-293: positions come from the DENSE enumeration of `patches`; the pixels written are the
    step-selected `patchcnn[cnt]` in the same order (overwrite)."""
    i_h, i_w = image_size[:2]
    p_h, p_w = patches.shape[1:3]
    img = np.zeros(image_size)
    n_h = i_h - p_h + 1
    n_w = i_w - p_w + 1
    cnt = 0
    for _, (i, j) in zip(patches, product(range(n_h), range(n_w))):
        if i % step == 0 and j % step == 0:
            img[i:i + p_h, j:j + p_w] = patchcnn[cnt]
            cnt += 1
    return img


def extract_patches_2dlocal(image, patches, patch_size, step=None):
    """imgpatch.py:295-337: select the dense patches whose (i, j) are multiples of `step`."""
    i_h, i_w = image.shape[:2]
    p_h, p_w = patch_size
    if p_h > i_h:
        raise ValueError("Height of the patch should be less than the height"
                         " of the image.")
    if p_w > i_w:
        raise ValueError("Width of the patch should be less than the width"
                         " of the image.")
    p_h, p_w = patches.shape[1:3]
    n_h = i_h - p_h + 1
    n_w = i_w - p_w + 1
    keep = [k for k, (i, j) in zip(range(len(patches)), product(range(n_h), range(n_w)))
            if i % step == 0 and j % step == 0]
    new_patch = np.zeros((len(keep), p_h, p_w, 3))
    for n, k in enumerate(keep):
        new_patch[n] = patches[k]
    return new_patch
