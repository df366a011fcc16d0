"""Drop-in mirror of the reference's img_utils.py hot-path functions, executed on the B200.

Reference: /root/reference/img_utils.py.  Same names, positional arguments, return types and error
behaviour; the Python double loops over numpy float64 are replaced by libsr100 gather/scatter kernels
(csrc/elementwise.cu) and are bit-exact for uint8-valued / float32-representable inputs.
Dropped on purpose: the reference's debug prints, its debug re-assembly and the hard-coded
imsave('/home/www/imgsuper/val_images/test.png') side effect (img_utils.py:661-674), and the
import-time os.makedirs of the dataset directory (img_utils.py:41-42; created lazily instead).
"""
from __future__ import print_function, division, absolute_import

import os

import numpy as np

# module constants of the reference (img_utils.py:21-39)
_image_scale_multiplier = 1
img_size = 256 * _image_scale_multiplier
stride = 16 * _image_scale_multiplier
input_path = r"input_images/"
validation_path = r"val_images/"
validation_set5_path = validation_path + "set5/"
validation_set14_path = validation_path + "set14/"
base_dataset_dir = os.path.expanduser("~") + "/Image Super Resolution Dataset/"
output_path = base_dataset_dir + "train_images/train/"
validation_output_path = base_dataset_dir + r"train_images/validation/"


def _ops():
    from sr100 import ops
    return ops


def extract_patches_Step(image, patch_size, step_patches=24):
    """img_utils.py:601-676.  image: (H, W, 3) array; returns (float64 (N, ph, pw, 3), (cnt_h, cnt_w)).
    Patch positions {x : 0 <= x < dim - p, x % step == 0} per axis; order w outer, h inner."""
    import torch
    ops = _ops()
    image = np.asarray(image)
    i_h, i_w = image.shape[:2]
    p_h, p_w = patch_size
    if p_h > i_h:
        raise ValueError("Height of the patch should be less than the height"
                         " of the image.")
    if p_w > i_w:
        raise ValueError("Width of the patch should be less than the width"
                         " of the image.")
    canvas = ops.to_device(image.reshape(i_h, i_w, -1), torch.float32)
    if canvas.shape[2] != 3:
        raise ValueError("extract_patches_Step expects 3 colour channels (img_utils.py:650)")
    patches, counts = ops.patch_gather_f32(canvas, (p_h, p_w), int(step_patches))
    return patches.cpu().numpy().astype(np.float64), counts


def rebuild_from_patches_Step(img_initial, patches, patch_size, tupleinit, scale, step_patches_ini=24):
    """img_utils.py:692-724.  patches: (N, ph*scale, pw*scale, 3); returns float64 (H*scale, W*scale, 3)
    with the 8-px border crop and last-writer-wins overlap of the reference loops."""
    import torch
    ops = _ops()
    i_h, i_w = np.asarray(img_initial).shape[:2]
    pd = ops.to_device(np.asarray(patches), torch.float32)
    out, _ = ops.patch_stitch(pd, tuple(tupleinit), tuple(patch_size), int(step_patches_ini), int(scale),
                              (i_h, i_w), mul=1.0, want_f32=True)
    return out.cpu().numpy().astype(np.float64)


# ------------------------------------------------------------------ data pipeline (feeds Model.fit_generator)
def _listdir_images(d):
    return sorted(f for f in os.listdir(d) if not f.startswith("."))


def image_count():
    """img_utils.py:126-128."""
    return len([name for name in os.listdir(output_path + "X/")])


def val_image_count():
    """img_utils.py:130-131."""
    return len([name for name in os.listdir(validation_output_path + "X/")])


def _imread_rgb(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"))


def _index_generator(N, batch_size=32, shuffle=True, seed=None):
    """img_utils.py:374-398: reshuffle at each epoch start, last batch short; yields
    (index_array, current_index, current_batch_size)."""
    batch_index = 0
    total_batches_seen = 0
    while 1:
        if seed is not None:
            np.random.seed(seed + total_batches_seen)
        if batch_index == 0:
            index_array = np.arange(N)
            if shuffle:
                index_array = np.random.permutation(N)
        current_index = (batch_index * batch_size) % N
        if N >= current_index + batch_size:
            current_batch_size = batch_size
            batch_index += 1
        else:
            current_batch_size = N - current_index
            batch_index = 0
        total_batches_seen += 1
        yield (index_array[current_index: current_index + current_batch_size],
               current_index, current_batch_size)


def image_generator(directory, scale_factor=2, target_shape=None, channels=3, small_train_images=False,
                    shuffle=True, batch_size=32, seed=None, device_resident=False):
    """img_utils.py:290-372: yields (batch_x, batch_y) float NHWC in [0,1] read from <directory>/X and
    <directory>/y (same file names).

    device_resident=True (not in the reference): the files are decoded once into HBM and every batch is one gather
    launch (sr100.dataset.DeviceDataset); the batches are device float32 tensors with the same values, in the same
    order, as the host arrays below."""
    if device_resident:
        from sr100.dataset import DeviceDataset
        ds = DeviceDataset(directory)
        print("Found %d images." % len(ds))
        for batch in ds.generator(batch_size, shuffle, seed):
            yield batch
    file_names = [f for f in _listdir_images(directory + "X/")]
    X_filenames = [os.path.join(directory, "X", f) for f in file_names]
    y_filenames = [os.path.join(directory, "y", f) for f in file_names]
    nb_images = len(file_names)
    print("Found %d images." % nb_images)
    index_generator = _index_generator(nb_images, batch_size, shuffle, seed)
    while 1:
        index_array, current_index, current_batch_size = next(index_generator)
        batch_x, batch_y = None, None
        for i, j in enumerate(index_array):
            img = _imread_rgb(X_filenames[j]).astype("float32") / 255.
            if batch_x is None:
                batch_x = np.zeros((current_batch_size,) + img.shape)
            batch_x[i] = img
            img = _imread_rgb(y_filenames[j]).astype("float32") / 255.
            if batch_y is None:
                batch_y = np.zeros((current_batch_size,) + img.shape)
            batch_y[i] = img
        yield (batch_x, batch_y)
