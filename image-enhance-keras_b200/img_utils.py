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


# ------------------------------------------------------------------ dataset preparation (img_utils.py:44-123)
def transform_images(directory, output_directory, scaling_factor=2, max_nb_images=-1, true_upscale=False):
    """img_utils.py:44-123: every image of `directory` -> 256 x 256 (bilinear) -> sharpen -> 256 sub-images of
    16*scaling_factor px; each is saved as ground truth under <out>/y/ and, blurred (sigma 0.5) + bicubic-shrunk to
    16 px (+ enlarged again unless true_upscale), under <out>/X/ as '<image index>_<sample index>.png'.
    The pixel work runs on the device (sr100.dataprep) and reproduces the reference's files bit for bit; decode and
    PNG encode stay on the host (PIL).  Kept as in the reference: the odd max_nb_images assertion, the early stop
    `index >= max_nb_images` (one image fewer than asked for), and the y samples being contrast-stretched by imsave."""
    import time
    import torch
    from PIL import Image
    from sr100 import dataprep
    index = 1
    for sub in ("X/", "y/"):
        if not os.path.exists(output_directory + sub):
            os.makedirs(output_directory + sub)
    nb_images = len([name for name in os.listdir(directory)])
    if max_nb_images != -1:
        print("Transforming %d images." % max_nb_images)
    else:
        assert max_nb_images <= nb_images, "Max number of images must be less than number of images in path"
        print("Transforming %d images." % (nb_images))
    if nb_images == 0:
        print("Extract the training images or images from imageset_91.zip (found in the releases of the project) "
              "into a directory with the name 'input_images'")
        print("Extract the validation images or images from set5_validation.zip (found in the releases of the project) "
              "into a directory with the name 'val_images'")
        exit()
    for file in os.listdir(directory):
        t1 = time.time()
        img = torch.from_numpy(np.array(_imread_rgb(directory + file), dtype=np.uint8)).cuda()
        y, x = dataprep.transform_image_device(img, scaling_factor, true_upscale)
        y, x = y.cpu().numpy(), x.cpu().numpy()
        for i in range(y.shape[0]):
            Image.fromarray(y[i]).save(output_directory + "/y/" + "%d_%d.png" % (index, i + 1))
            Image.fromarray(x[i]).save(output_directory + "/X/" + "%d_%d.png" % (index, i + 1))
        print("Finished image %d in time %0.2f seconds. (%s)" % (index, time.time() - t1, file))
        index += 1
        if max_nb_images > 0 and index >= max_nb_images:
            print("Transformed maximum number of images. ")
            break
    print("Images transformed. Saved at directory : %s" % (output_directory))


def subimage_generator(img, stride, patch_size, nb_hr_images):
    """img_utils.py:134-140."""
    for _ in range(nb_hr_images):
        for x in range(0, img_size - patch_size, stride):
            for y in range(0, img_size - patch_size, stride):
                yield img[x: x + patch_size, y: y + patch_size, :]


# ------------------------------------------------------------------ data pipeline (feeds Model.fit_generator)
def _listdir_images(d):
    return sorted(f for f in os.listdir(d) if not f.startswith("."))


def image_count():
    """img_utils.py:126-128."""
    return len(_listdir_images(output_path + "X/"))      # the same filter as the generators (no dotfiles)


def val_image_count():
    """img_utils.py:130-131."""
    return len(_listdir_images(validation_output_path + "X/"))


def _imread_rgb(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"))


def _index_generator(N, batch_size=32, shuffle=True, seed=None):
    """img_utils.py:374-398 (the Keras-1 index generator).  Yields (indices, start, count) forever.  The order is a
    contract (batches must be bit-identical to the reference's for a given seed), so its quirks are kept:
      * the global numpy RNG is re-seeded with seed + <batches yielded so far> before EVERY batch;
      * a new order (permutation or arange) is drawn only when the pass counter is 0, and the counter is reset only by
        a SHORT batch -- with N a multiple of batch_size the first order is reused for ever;
      * a batch that would run past N is cut short (start + batch_size > N) and ends the pass.
    Pinned to the reference's own function by tests/test_data_pipeline.py (tests/golden/generator_ref.npz)."""
    order, in_pass, yielded = None, 0, 0
    while True:
        if seed is not None:
            np.random.seed(seed + yielded)
        if in_pass == 0:
            order = np.random.permutation(N) if shuffle else np.arange(N)
        start = (in_pass * batch_size) % N
        count = batch_size if start + batch_size <= N else N - start
        in_pass = in_pass + 1 if count == batch_size else 0
        yielded += 1
        yield order[start:start + count], start, count


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


# ------------------------------------------------------------------ alternative-tiler helpers (img_utils.py:159-287, 442-599)
# Used by BaseSuperResolutionModel.upscalePatch / upscale(mode='patch').  The reference materialises every dense patch
# in host memory and loops over them in Python; the models' own paths (sr100.alt_tilers) never do.  These mirrors keep
# the function API: the extraction helpers are strided numpy views / index glue like the reference's, the two
# averaging reconstructions run on the device (sr_patch_average_*: float64 sums in the reference's patch order).
def imresize(arr, size, interp='bilinear', mode=None):
    """scipy.misc.imresize (scipy < 1.3; `from scipy.misc import imresize`, img_utils.py:5): toimage (bytescale for
    non-uint8 input) -> PIL resize -> uint8 array.  size: int (percent), float (fraction) or (rows, cols)."""
    from PIL import Image
    arr = np.asarray(arr)
    if arr.dtype != np.uint8:                     # scipy.misc.bytescale
        cmin, cmax = arr.min(), arr.max()
        cscale = cmax - cmin
        if cscale < 0:
            raise ValueError("`cmax` should be larger than `cmin`.")
        elif cscale == 0:
            cscale = 1
        arr = (((arr - cmin) * (255.0 / cscale)).clip(0, 255) + 0.5).astype(np.uint8)
    im = Image.fromarray(arr, mode=mode) if mode is not None else Image.fromarray(arr)
    if isinstance(size, (int, np.integer)):
        size = tuple((np.array(im.size) * (size / 100.0)).astype(int))
    elif isinstance(size, (float, np.floating)):
        size = tuple((np.array(im.size) * size).astype(int))
    else:
        size = (size[1], size[0])
    func = {'nearest': 0, 'lanczos': 1, 'bilinear': 2, 'bicubic': 3, 'cubic': 3}
    return np.asarray(im.resize(size, resample=func[interp]))


def extract_patches_2dv2(image, patch_size, max_patches=None, random_state=None):
    """img_utils.py:561-599: sklearn extract_patches_2d of image.astype('uint8')."""
    import imgpatch
    return imgpatch.extract_patches_2d(np.asarray(image).astype('uint8'), patch_size, max_patches, random_state)


def make_patches(x, scale, patch_size, upscale=True, verbose=1):
    """img_utils.py:159-172: every dense patch (sklearn extract_patches_2d); `scale` / `upscale` are ignored there."""
    import imgpatch
    return imgpatch.extract_patches_2d(np.asarray(x), (patch_size, patch_size))


def make_patchesOrig(x, scale, patch_size, upscale=False, verbose=1):
    """img_utils.py:174-180."""
    height, width = x.shape[:2]
    if upscale:
        x = imresize(x, (height * scale, width * scale))
    return extract_patches_2dv2(x, (patch_size, patch_size))


def make_patchesStep(x, scale, patch_size, upscale=False, extraction_step=24, verbose=1):
    """img_utils.py:182-187."""
    height, width = x.shape[:2]
    if upscale:
        x = imresize(x, (height * scale, width * scale))
    return extract_patches_Step(x, (patch_size, patch_size), extraction_step)


def extract_patches_2dlocal(image, patches, patch_size, step=None):
    """img_utils.py:513-556: the dense patches whose (i, j) are multiples of `step`, as float64."""
    i_h, i_w = image.shape[:2]
    p_h, p_w = patch_size
    if p_h > i_h:
        raise ValueError("Height of the patch should be less than the height"
                         " of the image.")
    if p_w > i_w:
        raise ValueError("Width of the patch should be less than the width"
                         " of the image.")
    p_h, p_w = patches.shape[1:3]
    n_h, n_w = i_h - p_h + 1, i_w - p_w + 1
    dense = np.asarray(patches)[:n_h * n_w].reshape((n_h, n_w) + patches.shape[1:])
    sel = dense[::step, ::step]
    out = np.zeros((sel.shape[0] * sel.shape[1], p_h, p_w, 3))
    out[...] = sel.reshape((-1,) + patches.shape[1:])
    return out


def _as_f32_patches(p):
    p = np.asarray(p)
    if p.ndim != 4 or p.shape[3] != 3:
        raise ValueError("expected patches of shape (n, p_h, p_w, 3), got %s" % (p.shape,))
    if p.shape[1] != p.shape[2]:
        raise ValueError("the device averaging kernel handles square patches, got %s" % (p.shape,))
    p32 = np.ascontiguousarray(p, dtype=np.float32)
    if p.dtype != np.float32 and not np.array_equal(p32, p):
        raise ValueError("patch values must be float32-representable (the network's outputs are)")
    return p32


def _average(patch_values, image_size, n_hw, step, pad, sklearn_count):
    import torch
    from sr100 import alt_tilers
    i_h, i_w = image_size[:2]
    n_h, n_w = n_hw
    P = patch_values.shape[1]
    cnt_h, cnt_w = (n_h - 1) // step + 1, (n_w - 1) // step + 1
    if patch_values.shape[0] < cnt_h * cnt_w:
        raise ValueError("need %d patches, got %d" % (cnt_h * cnt_w, patch_values.shape[0]))
    pd = torch.from_numpy(patch_values[:cnt_h * cnt_w]).cuda()
    edges = ((n_h - 1) // step if (n_h - 1) % step == 0 else -1, (n_w - 1) // step if (n_w - 1) % step == 0 else -1)
    _, f64 = alt_tilers.patch_average([(0, cnt_h, pd)], P, step, pad, cnt_h, cnt_w, (i_h, i_w), mul=1.0, want_f64=True,
                                      sklearn_count=sklearn_count, edges=edges)
    return f64.cpu().numpy()


def reconstruct_from_patches_2dlocal(patches, patchcnn, image_size, step=16):
    """img_utils.py:442-511: `patches` (the dense set) only gives the patch shape and the (i, j) enumeration; the
    values averaged are the step-selected `patchcnn`; interior patches contribute their [4, p-4) window; division by
    the count map (NaN where nothing contributes, as in the reference)."""
    i_h, i_w = image_size[:2]
    p_h, p_w = patches.shape[1:3]
    return _average(_as_f32_patches(patchcnn), image_size, (i_h - p_h + 1, i_w - p_w + 1), int(step), 4, False)


def reconstruct_from_patches_2dloc(patches, image_size):
    """img_utils.py:195-238: sklearn's averaging reconstruction (the code after its `return` is dead)."""
    i_h, i_w = image_size[:2]
    p = _as_f32_patches(patches)
    return _average(p, image_size, (i_h - p.shape[1] + 1, i_w - p.shape[2] + 1), 1, 0, True)


def combine_patches(in_patches, out_shape, scale):
    """img_utils.py:189-193: sklearn.feature_extraction.image.reconstruct_from_patches_2d(in_patches, out_shape)."""
    return reconstruct_from_patches_2dloc(in_patches, out_shape)


def subimage_build_patch_global(img, stride, patch_size, nb_hr_images):
    """img_utils.py:240-265 (the reference mixes the axes: y runs over range(width) but indexes rows)."""
    heightini, widthini = img.shape[:2]
    pos = [(y, x) for y in range(0, widthini, stride) for x in range(0, heightini, stride)
           if (x + patch_size) < widthini and (y + patch_size) < heightini]
    subimages = np.empty((len(pos), patch_size, patch_size, 3))
    for j, (y, x) in enumerate(pos):
        subimages[j] = img[y:y + patch_size, x:x + patch_size, :]
    return subimages


def subimage_combine_patches_global(imgtrue, patches, stride, patch_size, scale):
    """img_utils.py:268-287: overwrite the bicubic-upscaled image with the patches, same position rule."""
    heighttrue, widthtrue = imgtrue.shape[:2]
    img = imresize(imgtrue, (heighttrue * scale, widthtrue * scale), interp='bicubic').copy()
    heightini, widthini = img.shape[:2]
    j = 0
    for y in range(0, widthini, stride):
        for x in range(0, heightini, stride):
            if (x + patch_size) < widthini and (y + patch_size) < heightini:
                img[y:y + patch_size, x:x + patch_size, :] = patches[j, :, :, :]
                j += 1
    return img
