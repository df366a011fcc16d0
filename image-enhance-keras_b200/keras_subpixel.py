"""Drop-in mirror of the reference's keras_subpixel.py: `Subpixel(Conv2D)` (keras_subpixel.py:28-122).

A conv with r*r*filters outputs followed by the reference's _phase_shift (keras_subpixel.py:64-84), whose
index map is  out[b, Y, X, c] = conv[b, Y//r, X//r, c*r*r + (X%r)*r + (Y%r)]  (a pixel shuffle with the
two sub-pixel axes swapped).  Here the shuffle is not a separate pass: it is the store address of the
conv kernel's epilogue (the tcgen05 kernel's `shuffle_r` for 128-channel bf16 inputs, sr_conv2d_direct
otherwise; shuffle_order 0), so the r*r-channel tensor never round-trips HBM.  The layer is callable on numpy NHWC arrays or device tensors.
"""
import numpy as np


class Subpixel(object):
    def __init__(self, filters, kernel_size, r, padding='valid', data_format=None, strides=(1, 1),
                 activation=None, use_bias=True, kernel_initializer='glorot_uniform', bias_initializer='zeros',
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None,
                 kernel_constraint=None, bias_constraint=None, **kwargs):
        if isinstance(kernel_size, int):
            kernel_size = (kernel_size, kernel_size)
        if tuple(strides) != (1, 1):
            raise ValueError("Subpixel: only strides=(1,1) is implemented")
        if kernel_size[0] != kernel_size[1]:
            raise ValueError("Subpixel: only square kernels are implemented")
        if padding not in ('valid', 'same'):
            raise ValueError("Invalid border mode: " + str(padding))
        if activation not in (None, 'linear', 'relu'):
            raise ValueError("Subpixel: activation must be None, 'linear' or 'relu'")
        if data_format not in (None, 'channels_last'):
            raise ValueError("Subpixel: only channels_last is implemented")
        self.r = r
        self.filters = r * r * filters           # Conv2D filters (keras_subpixel.py:47)
        self.kernel_size = tuple(kernel_size)
        self.padding, self.activation, self.use_bias = padding, activation, use_bias
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        self.name = kwargs.get('name', 'subpixel')
        self.seed = kwargs.get('seed', None)
        self.kernel = None
        self.bias = None
        self.built = False

    def build(self, input_shape):
        cin = int(input_shape[-1])
        k = self.kernel_size[0]
        rng = np.random.default_rng(self.seed)
        limit = np.sqrt(6.0 / (k * k * cin + k * k * self.filters))   # glorot_uniform
        self.kernel = rng.uniform(-limit, limit, size=(k, k, cin, self.filters)).astype(np.float32)
        self.bias = np.zeros((self.filters,), dtype=np.float32) if self.use_bias else None
        self.built = True

    def get_weights(self):
        return [self.kernel] + ([self.bias] if self.use_bias else [])

    def set_weights(self, ws):
        self.kernel = np.ascontiguousarray(ws[0], dtype=np.float32)
        if self.use_bias:
            self.bias = np.ascontiguousarray(ws[1], dtype=np.float32)
        self.built = True

    def call(self, inputs):
        """keras_subpixel.py:109-110: _phase_shift(Conv2D.call(inputs)), fused."""
        import torch
        from sr100 import ops
        as_numpy = isinstance(inputs, np.ndarray)
        x = ops.to_device(inputs, torch.float32) if as_numpy else inputs.to(torch.float32).contiguous()
        if not self.built:
            self.build(x.shape)
        w = ops.to_device(self.kernel)
        b = ops.to_device(self.bias) if self.use_bias else None
        k = self.kernel_size[0]
        if (x.shape[-1] == 128 and self.padding == 'same' and k in (1, 3, 5) and 16 < self.filters <= 128
                and not isinstance(inputs, np.ndarray) and inputs.dtype == torch.bfloat16):
            # a 128-channel bf16 feature map (the width of the models.py stack): tensor-core conv, shuffle in its epilogue
            y = ops.conv2d_tc_shuffle(inputs.contiguous(), w, b, self.r, 0, relu=(self.activation == 'relu'))
        else:
            y = ops.conv2d_direct(x, w, b, same=(self.padding == 'same'), relu=(self.activation == 'relu'),
                                  shuffle_r=self.r, shuffle_order=0)
        return y.cpu().numpy() if as_numpy else y

    __call__ = call

    def compute_output_shape(self, input_shape):
        """keras_subpixel.py:112-114."""
        k = self.kernel_size[0]
        if self.padding == 'same':
            h, w = input_shape[1], input_shape[2]
        else:
            h = None if input_shape[1] is None else input_shape[1] - k + 1
            w = None if input_shape[2] is None else input_shape[2] - k + 1
        unshifted = (input_shape[0], h, w, self.filters)
        return (unshifted[0], None if h is None else self.r * unshifted[1], None if w is None else self.r * unshifted[2],
                int(unshifted[3] / (self.r * self.r)))

    def get_config(self):
        """keras_subpixel.py:116-122 ('rank' and 'dilation_rate' removed, filters divided by r*r, r added)."""
        return dict(name=self.name, filters=self.filters / (self.r * self.r), kernel_size=self.kernel_size,
                    strides=(1, 1), padding=self.padding, data_format='channels_last', activation=self.activation,
                    use_bias=self.use_bias, kernel_initializer=self.kernel_initializer,
                    bias_initializer=self.bias_initializer, r=self.r)
