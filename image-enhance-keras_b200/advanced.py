"""Drop-in mirror of the reference's advanced.py: HistoryCheckpoint (advanced.py:9-46), the depth_to_scale
functions (advanced.py:87-129), SubPixelUpscaling (advanced.py:135-159) and SubpixelConv2D
(advanced.py:173-199).  The shuffles run as libsr100 depth-to-space kernels; the three layers use three
different channel orderings (index maps derived from the reference code):

  depth_to_scale_tf / SubPixelUpscaling (TF backend):  ch = c*r*r + (X%r)*r + (Y%r)      (order 0)
  depth_to_scale_th (Theano, NCHW):                    ch = c*r*r + (Y%r)*r + (X%r)      (order 1)
  SubpixelConv2D -> tf.depth_to_space:                 ch = ((Y%r)*r + (X%r))*C + c      (order 2)
"""
import numpy as np


class HistoryCheckpoint(object):
    """advanced.py:9-46: records per-epoch logs and rewrites str(history) to `filename` every epoch."""

    def __init__(self, filename):
        self.filename = filename
        self.model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs={}):
        self.epoch = []
        self.history = {}

    def on_epoch_end(self, epoch, logs={}):
        self.epoch.append(epoch)
        for k, v in logs.items():
            if k not in self.history:
                self.history[k] = []
            self.history[k].append(v)
        with open(self.filename, "w") as f:
            f.write(str(self.history))


def _shuffle(x, r, order, channels_first=False):
    import torch
    from sr100 import ops
    as_numpy = isinstance(x, np.ndarray)
    t = ops.to_device(x, torch.float32) if as_numpy else x.to(torch.float32)
    if channels_first:
        t = t.permute(0, 2, 3, 1)
    y = ops.depth_to_space(t.contiguous(), r, order)
    if channels_first:
        y = y.permute(0, 3, 1, 2).contiguous()
    return y.cpu().numpy() if as_numpy else y


def depth_to_scale_th(input, scale, channels):
    """advanced.py:87-100 (NCHW): out[:, :, y::r, x::r] = input[:, r*y + x :: r*r]."""
    if input.shape[1] != channels * scale * scale:
        raise ValueError("depth_to_scale_th: input has %d channels, expected %d" % (input.shape[1], channels * scale * scale))
    # the Theano slicing input[:, r*y+x::r*r] is channel-minor: ch = c*r*r + y*r + x  -> order 1
    return _shuffle(input, scale, 1, channels_first=True)


def depth_to_scale_tf(input, scale, channels):
    """advanced.py:104-129 (NHWC): per colour group _phase_shift, groups concatenated on the channel axis
    (the reference hard-codes 3 groups, advanced.py:125)."""
    if channels > 1 and channels != 3:
        raise ValueError("depth_to_scale_tf splits the input into exactly 3 channel groups (advanced.py:125)")
    return _shuffle(input, scale, 0)


class SubPixelUpscaling(object):
    """advanced.py:135-159."""

    def __init__(self, r, channels, **kwargs):
        self.r = r
        self.channels = channels
        self.name = kwargs.get('name', 'subpixelupscaling')

    def build(self, input_shape):
        pass

    def call(self, x, mask=None):
        return depth_to_scale_tf(x, self.r, self.channels)

    __call__ = call

    def get_output_shape_for(self, input_shape):
        b, r, c, k = input_shape
        return (b, r * self.r, c * self.r, self.channels)

    compute_output_shape = get_output_shape_for


class _Lambda(object):
    def __init__(self, fn, output_shape, name):
        self.function, self._output_shape, self.name = fn, output_shape, name

    def __call__(self, x):
        return self.function(x)

    def compute_output_shape(self, input_shape):
        return self._output_shape(input_shape)


def SubpixelConv2D(input_shape, scale=4):
    """advanced.py:173-199: Lambda(tf.depth_to_space(x, scale), name='subpixel')."""

    def subpixel_shape(input_shape):
        dims = [input_shape[0], input_shape[1] * scale, input_shape[2] * scale, int(input_shape[3] / (scale ** 2))]
        return tuple(dims)

    def subpixel(x):
        return _shuffle(x, scale, 2)

    return _Lambda(subpixel, subpixel_shape, 'subpixel')
