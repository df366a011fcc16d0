"""Drop-in mirror of the reference's models.py for the x4 hot path (reference: /root/reference/models.py).

Same constructors, attributes, method names / argument meaning / error behaviour as the reference
classes; the Keras/TensorFlow graph is replaced by the sm_100a engine (sr100.engine) reached through
the libsr100 C ABI.  There is no CPU path: constructing a model's graph needs a B200.

Covered: psnr helpers (models.py:43-90), BaseSuperResolutionModel (:93-182), upscaleStepPatch
(:184-415, the CLI path), upscalePatch (:419-604) and upscale (:606-852, 'fast' and 'patch' modes) on the
alternative tilers of sr100.alt_tilers, evaluate (:159-163, 1519-1622), DifvdsrDouble (:1146-1270) with the
Lambda helpers it uses (:977-986, :1383-1399, :1451), and the Difvdsr4 (:992-1142) / Difvdsr (:1274-1357)
graphs (sr100.planenet).  _evaluate_denoise is out of scope (auto-encoder models this file does not define).
"""
from __future__ import print_function, division

import math
import os

import numpy as np

import img_utils
from advanced import HistoryCheckpoint, SubPixelUpscaling, SubpixelConv2D  # noqa: F401 (models.py:16)
from keras_subpixel import Subpixel  # noqa: F401 (models.py:17)

train_path = img_utils.output_path
validation_path = img_utils.validation_output_path
path_X = img_utils.output_path + "X/"
path_Y = img_utils.output_path + "y/"


def _mean_sq(a, b):
    from sr100 import ops
    a, b = np.asarray(a), np.asarray(b)
    return ops.sum_sq_diff(a, b) / a.size


def PSNRLoss(y_true, y_pred):
    """models.py:43-55: the reference returns K.mean(y_pred) (the PSNR formula after it is dead code)."""
    return float(np.mean(np.asarray(y_pred)))


def PSNRLossTest(y_true, y_pred):
    """models.py:57-69."""
    return -10. * np.log10(_mean_sq(y_pred, y_true))


def psnr(y_true, y_pred):
    """models.py:71-76."""
    assert y_true.shape == y_pred.shape, "Cannot calculate PSNR. Input shapes not same." \
                                         " y_true shape = %s, y_pred shape = %s" % (str(y_true.shape),
                                                                                   str(y_pred.shape))
    return -10. * np.log10(_mean_sq(y_pred, y_true))


def psnr2(img1, img2):
    """models.py:78-83."""
    mse = _mean_sq(img1, img2)
    if mse == 0:
        return 100
    return 20 * math.log10(255.0 / math.sqrt(mse))


def psnr3(img1, img2):
    """models.py:85-90 (keeps the reference's stray sqrt)."""
    mse = _mean_sq(img1, img2)
    if mse == 0:
        return 100
    return 10 * math.log10((255.0 ** 2) / math.sqrt(mse))


# ---- Lambda helpers used by DifvdsrDouble (tensor-in / tensor-out on device NHWC float32 tensors) ----
def resizeBlockLight09(my_input):
    """models.py:977-978: tf.scalar_mul(0.9, x) (fused into the conv epilogue inside the engine)."""
    return my_input * 0.9


def resizeBlockLight01(my_input):
    """models.py:985-986."""
    return my_input * 0.1


def resizeRes_outputshape(my_input_shape):
    """models.py:1451."""
    return my_input_shape


def resizeX4_outputshape(my_input_shape):
    """models.py:1383-1390."""
    shape = list(my_input_shape)
    size = [4 * int(s) for s in my_input_shape[1:3]]
    shape[1] = size[0]
    shape[2] = size[1]
    return tuple(shape)


resizeX4bil_outputshape = resizeX4_outputshape


def resizeX4bil(my_input):
    """models.py:1392-1399: tf.image.resize_bilinear(x, [4h, 4w]) with TF1 legacy sampling.
    Accepts a numpy NHWC array or a device tensor (C % 8 == 0); returns the same kind."""
    import torch
    from sr100 import ops
    if isinstance(my_input, np.ndarray):
        return ops.bilinear4(ops.to_device(my_input, torch.float32)).cpu().numpy()
    return ops.bilinear4(my_input)


class _Adam(object):
    """optimizers.Adam(lr, beta_1) record (models.py:1212); Keras-2 defaults for the rest."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, decay=0.):
        self.lr, self.beta_1, self.beta_2, self.epsilon, self.decay = lr, beta_1, beta_2, epsilon, decay


class _ModelCheckpoint(object):
    """keras.callbacks.ModelCheckpoint(filepath, save_weights_only=True, period=1) as used at
    models.py:141-142: saves every epoch to filepath.format(epoch=epoch+1, **logs)."""

    def __init__(self, filepath, monitor='val_loss', save_best_only=False, mode='auto', save_weights_only=True,
                 period=1):
        self.filepath, self.period, self.model = filepath, period, None

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        logs = dict(logs or {})
        logs.setdefault("val_acc", float("nan"))
        if (epoch + 1) % self.period == 0:
            path = self.filepath.format(epoch=epoch + 1, **logs)
            d = os.path.dirname(path)
            if d and not os.path.isdir(d):
                os.makedirs(d)
            self.model.save_weights(path, overwrite=True)


class BaseSuperResolutionModel(object):

    def __init__(self, model_name, scale_factor):
        """models.py:95-109."""
        self.model = None
        self.model_name = model_name
        self.scale_factor = scale_factor
        self.weight_path = None

        self.type_scale_type = "norm"  # Default = "norm" = 1. / 255
        self.type_requires_divisible_shape = False
        self.type_true_upscaling = False

        self.evaluation_func = None
        self.uses_learning_phase = False
        self._engine = None  # device-resident weights, shared by every create_model() of this object

    def create_model(self, height=32, width=32, channels=3, load_weights=False, batch_size=2):
        """models.py:111-129: validates the shape and returns the input shape (width-major like the reference)."""
        if self.type_requires_divisible_shape:
            assert height * img_utils._image_scale_multiplier % 4 == 0, "Height of the image must be divisible by 4"
            assert width * img_utils._image_scale_multiplier % 4 == 0, "Width of the image must be divisible by 4"
        shape = (width * img_utils._image_scale_multiplier, height * img_utils._image_scale_multiplier, channels)
        return shape

    def fit(self, batch_size=2, nb_epochs=100, save_history=True, history_fn="Model History.txt"):
        """models.py:131-157."""
        samples_per_epoch = img_utils.image_count()
        val_count = img_utils.val_image_count()
        if self.model is None:
            self.create_model(batch_size=batch_size)
        callback_list = [_ModelCheckpoint(self.weight_path, monitor='val_PSNRLoss', save_best_only=False,
                                          mode='max', save_weights_only=True, period=1)]
        if save_history:
            callback_list.append(HistoryCheckpoint(history_fn))
        print("Training model : %s" % (self.__class__.__name__))
        # SR100_DEVICE_DATASET=1 (default): decode the two directories once into HBM and assemble every batch on the
        # device (sr100.dataset); datasets that do not qualify (ragged image shapes, too large) use the reference's
        # per-batch host decode.  Either way the batches hold the same values in the same order.
        def gen(directory):
            kw = dict(scale_factor=self.scale_factor, small_train_images=self.type_true_upscaling,
                      batch_size=batch_size)
            if os.environ.get("SR100_DEVICE_DATASET", "1") != "0":
                try:
                    from sr100.dataset import DeviceDataset
                    ds = DeviceDataset(directory)
                    print("Found %d images." % len(ds))
                    return ds.generator(batch_size, True, None)
                except ValueError as e:
                    print("device-resident dataset not used (%s); decoding per batch on the host" % e)
            return img_utils.image_generator(directory, **kw)
        self.model.fit_generator(gen(train_path), steps_per_epoch=samples_per_epoch, epochs=nb_epochs,
                                 callbacks=callback_list, validation_data=gen(validation_path),
                                 validation_steps=val_count)
        return self.model

    def evaluate(self, validation_dir):
        """models.py:159-163."""
        if self.type_requires_divisible_shape:
            raise NotImplementedError("_evaluate_denoise (models.py:1625-1721) serves the denoising auto-encoders, "
                                      "none of which this file defines")
        _evaluate(self, validation_dir)

    def upVideo(self, imgObj, save_intermediate=False, return_image=False, suffix="scaled",
                patch_size=8, mode="patch", verbose=False):
        """models.py:165-182: whole-image forward.  Like the reference, the [0,1] network output is clipped and
        cast to uint8 WITHOUT the x255 rescale (models.py:181)."""
        img_width, img_height = imgObj.shape[0], imgObj.shape[1]
        images = np.expand_dims(imgObj, axis=0)
        img_conv = images.astype(np.float32) / 255.
        model = self.create_model(img_height, img_width, load_weights=True)
        result = model.predict(img_conv, batch_size=128, verbose=verbose)
        result = result[0, :, :, :]
        return np.clip(result, 0, 255).astype('uint8')

    def upscaleStepPatch(self, img_path, save_intermediate=False, return_image=False, suffix="scaled",
                         patch_size=256, scalemulti=4, step_patch=64, mode="patch", verbose=True):
        """models.py:184-415.  The whole chain (zero-pad canvas, 96/64 patch gather, /255, conv stack, x255,
        stitch with the 8-px crop, clip -> uint8 truncation) runs on the device; only the uint8 image goes in
        and the uint8 canvas comes back.  `step_patch` is ignored like in the reference (forced to 64, :248)."""
        import torch
        from PIL import Image
        from sr100 import ops

        path = os.path.splitext(img_path)
        filename = path[0] + "_" + suffix + "(%dx)" % (self.scale_factor) + path[1]

        true_img = np.asarray(Image.open(img_path).convert("RGB"))       # imread(img_path, mode='RGB'), :212
        orig_height, orig_width = true_img.shape[0], true_img.shape[1]
        step_patch = 64                                                   # :248
        if mode == 'patch':
            canvas_h, canvas_w = ops.canvas_size(orig_height, orig_width, patch_size, step_patch)  # :225-256
        else:
            canvas_h, canvas_w = orig_height + patch_size, orig_width + patch_size
        if verbose:
            print("Old Size : ", (canvas_h, canvas_w, 3))
        img_dev = ops.to_device(true_img, torch.uint8)

        if mode == 'patch':
            if verbose or save_intermediate:
                cnt = (ops.patch_count(canvas_h, patch_size, step_patch), ops.patch_count(canvas_w, patch_size, step_patch))
            if verbose:
                print("Number of patches = %d, Patch Shape = (%d, %d)" % (cnt[0] * cnt[1], patch_size, patch_size))
            if save_intermediate:
                fn = path[0] + "_intermediate_" + path[1]
                first = true_img[:patch_size, :patch_size].astype(np.float32)         # imsave(fn, images[0]), :329
                first = np.pad(first, ((0, patch_size - first.shape[0]), (0, patch_size - first.shape[1]), (0, 0)))
                lo, hi = first.min(), first.max()
                sc = 255.0 / (hi - lo) if hi > lo else 1.0
                Image.fromarray(((first - lo) * sc + 0.5).astype(np.uint8)).save(fn)
            model = self.create_model(patch_size, patch_size, load_weights=True)       # :338
            # :272-:391 on the device.  return_image hands back the whole canvas (:405-407) and therefore runs every
            # tile in full; the file output is the 4H x 4W crop (:412), for which tiles / patch regions that cannot
            # reach it are not computed (bit-identical pixels, see Engine.upscale_images_device)
            result_u8 = model.engine.upscale_images_device([img_dev], patch=patch_size, step=step_patch,
                                                           scale=scalemulti, full_canvas=bool(return_image))[0]
        else:
            canvas = torch.zeros(1, canvas_h, canvas_w, 3, device=img_dev.device, dtype=torch.float32)
            canvas[0, :orig_height, :orig_width] = img_dev.to(torch.float32) / 255.0
            model = self.create_model(canvas_h, canvas_w, load_weights=True)
            result = model.engine.forward_device(canvas)
            result_u8 = torch.clamp(result[0] * 255.0, 0, 255).to(torch.uint8)
        result = result_u8.cpu().numpy()
        if return_image:
            return result                                                             # uncropped canvas, :405-407
        if verbose:
            print("Saving image.")
        outresult = result[0:orig_height * scalemulti, 0:orig_width * scalemulti]      # :412 (no-op in patch mode)
        Image.fromarray(outresult).save(filename)                                     # :415

    def upscale_arrays(self, images, patch_size=96, scalemulti=4, return_canvas=False):
        """Batched in-memory form of upscaleStepPatch(mode='patch') (no file I/O): list of uint8 (H,W,3) host
        arrays -> list of uint8 (4H,4W,3) host arrays.  Tiles of all images share one pass over the conv stack.
        The timed end-to-end path of bench.py: host -> device copies in, device -> host copies out."""
        import torch
        images = [np.asarray(im) for im in images]
        for im in images:
            if im.ndim != 3 or im.shape[2] != 3 or im.dtype != np.uint8 or im.shape[0] < 1 or im.shape[1] < 1:
                raise ValueError("upscale_arrays expects uint8 arrays of shape (H, W, 3), got %s %s" % (im.shape, im.dtype))
        if not images:
            return []
        model = self.model if self.model is not None else self.create_model(patch_size, patch_size, load_weights=False)
        eng = model.engine
        dev = [torch.from_numpy(np.ascontiguousarray(im)).pin_memory().to(eng.device, non_blocking=True)
               for im in images]
        from sr100 import ops
        # return_canvas: every tile in full, uncropped canvases; default: the 4H x 4W images (:412), stitched directly
        canv = eng.upscale_images_device(dev, patch=patch_size, step=64, scale=scalemulti, full_canvas=return_canvas)
        pinned = [ops.to_host_pinned(c) for c in canv]
        torch.cuda.current_stream().synchronize()
        return [t.numpy() for t in pinned]

    def upscalePatch(self, img_path, save_intermediate=False, return_image=False, suffix="scaled",
                     patch_size=32, scalemulti=4, mode="patch", verbose=True):
        """models.py:419-604: "enhance at the same size".  The image is zero-padded to multiples of 4 (both sides
        bumped when either is off, :465-470), every 4th `patch_size` patch is shrunk x4 with
        scipy.misc.imresize(..., 'bicubic') (bytescale + Pillow bicubic, :487-490), run through the network (x4,
        back to patch_size) and the outputs are averaged by img_utils.reconstruct_from_patches_2dlocal (:556);
        clip -> uint8, crop to the original size (:575-577).  All of it on the device (sr100.alt_tilers)."""
        import torch
        from PIL import Image
        from sr100 import alt_tilers, ops
        path = os.path.splitext(img_path)
        filename = path[0] + "_" + suffix + "(%dx)" % (self.scale_factor) + path[1]
        true_img = np.asarray(Image.open(img_path).convert("RGB"))
        orig_height, orig_width = true_img.shape[0], true_img.shape[1]
        init_height, init_width = orig_height, orig_width
        if verbose:
            print("Old Size : ", true_img.shape)
            print("New Size : (%d, %d, 3)" % (init_height, init_width))
        if mode != 'patch':
            # the reference reads img_height before assigning it on this branch (models.py:501)
            raise UnboundLocalError("local variable 'img_height' referenced before assignment")
        if int(scalemulti) != 4:
            raise ValueError("upscalePatch shrinks every patch by scalemulti and the network enlarges it x4: "
                             "scalemulti must be 4")
        step_patch = 4
        if init_width % step_patch != 0 or init_height % step_patch != 0:
            new_w = int((init_width / step_patch) + 1) * step_patch
            new_h = int((init_height / step_patch) + 1) * step_patch
            new_img = np.zeros((new_h, new_w, 3), dtype=np.uint8)
            new_img[0:init_height, 0:init_width] = true_img
            true_img = new_img
            init_height, init_width = new_h, new_w
        if patch_size % 4 != 0:
            raise ValueError("patch_size must be a multiple of 4 (each patch is shrunk to patch_size/4)")
        q = patch_size // 4
        img_dev = ops.to_device(true_img, torch.uint8)
        if verbose:
            cnt = ((init_height - patch_size) // 4 + 1) * ((init_width - patch_size) // 4 + 1)
            print("Number of patches = %d, Patch Shape = (%d, %d)" % (cnt, q, q))
        if save_intermediate:                       # imsave(fn, images[0]) (:515): the first shrunk patch
            first = alt_tilers.patch_down4(img_dev, patch_size, 4, True, 0, 1)[0]
            Image.fromarray((first * 255.0).round().to(torch.uint8).cpu().numpy()).save(path[0] + "_intermediate_" + path[1])
        model = self.create_model(q, q, load_weights=True)
        if verbose:
            print("Model loaded.")
        result = alt_tilers.enhance_image_device(model.engine, img_dev, patch_size, 4, stretch=True, pad=4)
        result = result[0:orig_height, 0:orig_width].contiguous().cpu().numpy()
        if return_image:
            return result
        if verbose:
            print("Saving image.")
        Image.fromarray(result).save(filename)

    def upscale(self, img_path, save_intermediate=False, return_image=False, suffix="scaled",
                patch_size=32, mode="patch", verbose=True):
        """models.py:606-852.  mode='fast' (the whole image through the network, :681-758, :773-848): the image is
        bicubic-resized to its own size (`imresize(true_img, (img_width, img_height))` with
        __match_autoencoder_size returning the input size for this model, :855-889 -- a PIL no-op), written as the
        `_A<suffix>` side file (:755), run through model.predict, x255, clipped to uint8 and saved.
        mode='patch': dense sklearn patches of the x4-bicubic image (:645-680), see the branch below."""
        from PIL import Image
        path = os.path.splitext(img_path)
        filename = path[0] + "_" + suffix + "(%dx)" % (self.scale_factor) + path[1]
        filenameM = path[0] + "_A" + suffix + "(%dx)" % (self.scale_factor) + path[1]
        true_img = np.asarray(Image.open(img_path).convert("RGB"))
        rows, cols = true_img.shape[0], true_img.shape[1]      # the reference calls these init_width, init_height
        if verbose:
            print("Old Size : ", true_img.shape)
            print("New Size : (%d, %d, 3)" % (cols * int(self.scale_factor), rows * int(self.scale_factor)))
        if mode == "patch" and self.type_true_upscaling:
            mode = 'fast'
            print("Patch mode does not work with True Upscaling models yet. Defaulting to mode='fast'")
        if mode == 'patch':
            # :645-680, 758-790: bicubic x4 of the whole image (PIL, as scipy.misc.imresize does), the `_A<suffix>` side
            # file, then every dense patch_size patch shrunk x4 -> network -> sklearn averaging (combine_patches)
            import torch
            from sr100 import alt_tilers, ops
            if patch_size % 4 != 0:
                raise ValueError("patch_size must be a multiple of 4 (each patch is shrunk to patch_size/4)")
            big = np.asarray(Image.fromarray(true_img).resize((cols * 4, rows * 4), Image.BICUBIC))
            Image.fromarray(big).save(filenameM)                                        # :657
            if verbose:
                print("Number of patches = %d, Patch Shape = (%d, %d)"
                      % ((big.shape[0] - patch_size + 1) * (big.shape[1] - patch_size + 1), patch_size // 4, patch_size // 4))
            big_dev = ops.to_device(big, torch.uint8)
            if save_intermediate:
                first = alt_tilers.patch_down4(big_dev, patch_size, 1, False, 0, 1)[0]
                Image.fromarray((first * 255.0).round().to(torch.uint8).cpu().numpy()).save(path[0] + "_intermediate_" + path[1])
            model = self.create_model(patch_size // 4, patch_size // 4, load_weights=True)
            if verbose:
                print("Model loaded.")
            result = alt_tilers.enhance_image_device(model.engine, big_dev, patch_size, 1, stretch=False, pad=0)
            result = result.cpu().numpy()
            if return_image:
                return result
            if verbose:
                print("Saving image.")
            Image.fromarray(result).save(filename)
            return
        sf = int(self.scale_factor)
        img_height, img_width = cols * sf, rows * sf            # __match_autoencoder_size, not AE / not true upscaling
        if (img_width, img_height) == (rows, cols):
            images = true_img                                   # PIL resize to the same size returns a copy
        else:
            images = np.asarray(Image.fromarray(true_img).resize((img_height, img_width), Image.BICUBIC))
        Image.fromarray(images).save(filenameM)                  # imsave(filenameM, images), :755
        if save_intermediate:
            if verbose:
                print("Saving intermediate image.")
            Image.fromarray(images).save(path[0] + "_intermediate_" + path[1])
        img_conv = np.expand_dims(images, axis=0).astype(np.float32) / 255.
        model = self.create_model(img_height, img_width, load_weights=True)
        if verbose:
            print("Model loaded.")
        result = model.predict(img_conv, batch_size=10, verbose=verbose)
        if verbose:
            print("De-processing images.")
        result = result.astype(np.float32) * 255.
        result = np.clip(result[0], 0, 255).astype('uint8')
        if verbose:
            print("\nCompleted De-processing image.")
        if return_image:
            return result
        if verbose:
            print("Saving image.")
        Image.fromarray(result).save(filename)


class DifvdsrDouble(BaseSuperResolutionModel):
    """models.py:1146-1270: 1x1 head, 16 5/3 residual blocks, 6 light blocks, bilinear x4, 2 5/3 blocks at HR,
    3x3 tail; 86 convs, 21,838,211 parameters."""

    weights_file = "weights_Double/weights025-17-0.93.h5"   # models.py:1217
    # conv operand precision of the engine: None -> $SR100_PRECISION or "bf16"; "tf32" = the accuracy mode (fp32
    # tensors, tf32 MMAs, ~1e-4 of the reference's fp32 graph at half the tensor rate); inference only
    precision = None

    def __init__(self, scale_factor):
        super(DifvdsrDouble, self).__init__("Image ScaleGen", scale_factor)
        self.weight_path = "weights_Double/weights025-{epoch:02d}-{val_acc:.2f}.h5"   # models.py:1155
        self._loaded_from = None

    def create_model(self, height=32, width=32, channels=3, load_weights=False, batch_size=128):
        """models.py:1159-1222.  Returns the Keras-like Model (sr100.kmodel.Model) for (height, width) inputs."""
        from sr100.kmodel import Model
        from sr100.engine import Engine
        init = super(DifvdsrDouble, self).create_model(height, width, channels, load_weights, batch_size)
        if channels != 3:
            raise ValueError("DifvdsrDouble is a 3-channel model (models.py:1177)")
        if self._engine is None:
            self._engine = Engine(precision=self.precision)   # glorot_uniform kernels, zero biases (Keras defaults)
        model = Model(init, engine=self._engine)
        model.compile(optimizer=_Adam(1e-4, 0.9), loss='mse', metrics=['accuracy'])   # :1212-1213
        if load_weights:
            wpath = os.environ.get("SR100_WEIGHTS", self.weights_file)
            if self._loaded_from != wpath:               # the reference reloads per image (:338); once is enough
                model.load_weights(wpath)                # :1217-1218
                self._loaded_from = wpath
        self.model = model
        return model

    def fit(self, batch_size=10, nb_epochs=100, save_history=False, history_fn="ScaleGen History.txt"):
        """models.py:1225-1226."""
        return super(DifvdsrDouble, self).fit(batch_size, nb_epochs, save_history, history_fn)


class _PlaneNetModel(BaseSuperResolutionModel):
    """Shared create_model of the two older graphs (sr100.planenet.PlaneNet engines)."""
    arch = None
    weights_file = None
    force_load = False

    def create_model(self, height=32, width=32, channels=3, load_weights=False, batch_size=128):
        from sr100.kmodel import Model
        from sr100.planenet import PlaneNet
        init = BaseSuperResolutionModel.create_model(self, height, width, channels, load_weights, batch_size)
        if channels != 3:
            raise ValueError("%s is a 3-channel model" % self.__class__.__name__)
        if self._engine is None:
            self._engine = PlaneNet(self.arch)           # glorot_uniform kernels, zero biases (Keras defaults)
        model = Model(init, engine=self._engine)
        model.compile(optimizer=_Adam(1e-4, 0.9), loss='mse', metrics=['accuracy'])
        if load_weights or self.force_load:
            wpath = os.environ.get("SR100_WEIGHTS", self.weights_file)
            if getattr(self, "_loaded_from", None) != wpath:
                model.load_weights(wpath)
                self._loaded_from = wpath
        self.model = model
        return model


class Difvdsr4(_PlaneNetModel):
    """models.py:992-1142: 256 channels; 1x1 head, 6 LeakyReLU(0.001) light blocks, bilinear x2, 20 light blocks with
    a skip over them, bilinear x2, 6 light blocks, 3x3 tail; 66 convs."""
    arch = "difvdsr4"
    weights_file = "weights_Difvdsr2scale/0.1/weights025-18-0.94.h5"       # models.py:1066

    def __init__(self, scale_factor):
        super(Difvdsr4, self).__init__("Image ScaleGen", scale_factor)
        self.weight_path = "weights_Difvdsr2scale/weights-{epoch:02d}-{val_acc:.2f}.h5"   # models.py:1001

    def create_model(self, height=24, width=24, channels=3, load_weights=False, batch_size=1):
        """models.py:1006-1076."""
        return super(Difvdsr4, self).create_model(height, width, channels, load_weights, batch_size)

    def fit(self, batch_size=2, nb_epochs=100, save_history=False, history_fn="ScaleGen History.txt"):
        """models.py:1079-1080."""
        return super(Difvdsr4, self).fit(batch_size, nb_epochs, save_history, history_fn)


class Difvdsr(_PlaneNetModel):
    """models.py:1274-1357: 192 channels, same resolution; 3x3 head, 32 difference blocks (Subtract, LeakyReLU(0.2),
    3-way Add, x0.1), 3x3 tail; 130 convs.  Like the reference it always loads its weight file (models.py:1322)."""
    arch = "difvdsr"
    weights_file = "weights_Difvdsr/weights-23-0.96.h5"                    # models.py:1323
    force_load = True

    def __init__(self, scale_factor):
        super(Difvdsr, self).__init__("Image ScaleGen", scale_factor)
        self.weight_path = "weights_Difvdsr/weights-{epoch:02d}-{val_acc:.2f}.h5"         # models.py:1283

    def create_model(self, height=64, width=64, channels=3, load_weights=False, batch_size=128):
        """models.py:1288-1329."""
        return super(Difvdsr, self).create_model(height, width, channels, load_weights, batch_size)

    def fit(self, batch_size=8, nb_epochs=100, save_history=False, history_fn="ScaleGen History.txt"):
        """models.py:1332-1333."""
        return super(Difvdsr, self).fit(batch_size, nb_epochs, save_history, history_fn)


def _evaluate(sr_model, validation_dir, scale_pred=False):
    """models.py:1519-1622, line by line: every image of <validation_dir>/set5 and /set14 is read, scaled to [0,1],
    bicubic-"resized" to its own size twice through scipy.misc.imresize (which byte-scales the float image back to
    uint8 0..255 - the network is then fed 0..255 values, as in the reference), predicted, and compared with
    models.psnr; the prediction is written to val_predict/.  Like the reference it only makes sense for same-size
    models (Difvdsr): for a x4 model psnr() receives arrays of different shapes and numpy raises."""
    import time
    from PIL import Image
    print("Validating %s model" % sr_model.model_name)
    predict_path = "val_predict/"
    if not os.path.exists(predict_path):
        os.makedirs(predict_path)
    for val_dir in [validation_dir + "set5/", validation_dir + "set14/"]:
        image_fns = [name for name in os.listdir(val_dir)]
        nb_images = len(image_fns)
        print("Validating %d images from path %s" % (nb_images, val_dir))
        total_psnr = 0.0
        for impath in os.listdir(val_dir):
            t1 = time.time()
            y = np.asarray(Image.open(val_dir + impath).convert("RGB"))
            width, height, _ = y.shape                       # (rows, cols): the reference's naming
            y = y.astype('float32')
            sf = sr_model.scale_factor
            x_width = width if not sr_model.type_true_upscaling else width // sf
            x_height = height if not sr_model.type_true_upscaling else height // sf
            x_temp = y.copy()
            if sr_model.type_scale_type == "tanh":
                x_temp = (x_temp - 127.5) / 127.5
                y = (y - 127.5) / 127.5
            else:
                x_temp /= 255.
                y /= 255.
            y = np.expand_dims(y, axis=0)
            img = img_utils.imresize(x_temp, (x_width, x_height), interp='bicubic')
            if not sr_model.type_true_upscaling:
                img = img_utils.imresize(img, (x_width, x_height), interp='bicubic')
            x = np.expand_dims(img, axis=0)
            model = sr_model.create_model(x_height, x_width, load_weights=True)     # per-shape model (:1524)
            y_pred = model.predict(x.astype(np.float32), batch_size=1)[0]
            if scale_pred:
                y_pred = (y_pred + 1) * 127.5 if sr_model.type_scale_type == "tanh" else y_pred * 255.
            if sr_model.type_scale_type == 'tanh':
                y = (y + 1) / 2
            psnr_val = psnr(y[0], np.clip(y_pred, 0, 255) / 255)
            total_psnr += psnr_val
            t2 = time.time()
            print("Validated image : %s, Time required : %0.2f, PSNR value : %0.4f" % (impath, t2 - t1, psnr_val))
            generated_path = predict_path + "%s_%s_generated.png" % (sr_model.model_name, os.path.splitext(impath)[0])
            Image.fromarray(np.clip(y_pred, 0, 255).astype('uint8')).save(generated_path)
        print("Average PRNS value of validation images = %00.4f \n" % (total_psnr / nb_images))
