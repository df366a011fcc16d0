"""CPU restatement of the reference's float path (DifvdsrDouble, loss, optimizer).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product (image-enhance-keras_b200/) never does.

PARITY UNPINNED for this file: the reference graph runs on Keras 2.x / TensorFlow 1.x, neither of
which is installable in the authoring container (no network), and the reference ships no tests or
golden vectors for the conv stack (trained weights are an external download).  The restatement
below follows the reference source line by line plus the documented library semantics:

  * graph:      models.py:1159-1222 (create_model), :1231-1245 (_residual_block_light),
                :1248-1270 (_residual_block_light53); constants 0.9 / 0.1 from
                resizeBlockLight09 / resizeBlockLight01 (models.py:977-986).
  * Conv2D:     Keras cross-correlation, HWIO kernel, padding='same' (symmetric zero pad (k-1)/2 for
                odd k, stride 1), bias add, optional ReLU; glorot_uniform init, zero bias.
  * upsample:   tf.image.resize_bilinear(x, [4h, 4w]) (models.py:1392-1399), TF1 default
                align_corners=False with LEGACY sampling (no half-pixel centres).
  * scaling:    models.py:336 (/255.), :351 (*255.), :391 (clip -> uint8 truncation).
  * training:   loss 'mse', Adam(lr=1e-4, beta_1=0.9) with Keras-2 defaults beta_2=0.999,
                epsilon=1e-7, no decay (models.py:1203-1213).
Layer names follow Keras auto-naming in creation order: 'level1', 'conv2d_1' ... 'conv2d_85'
(k3,k5,k5,k3 inside a 5/3 block; k3,k3 inside a light block; tail last).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

NUMK = 128


def layer_specs():
    """[(name, ksize, cin, cout)] in Keras creation order (models.py:1177-1199)."""
    specs = [("level1", 1, 3, NUMK)]
    n = 0

    def add(k, cin=NUMK, cout=NUMK):
        nonlocal n
        n += 1
        specs.append(("conv2d_%d" % n, k, cin, cout))

    for _ in range(16):          # models.py:1182-1184, block body :1253-1259
        add(3); add(5); add(5); add(3)
    for _ in range(6):           # models.py:1188-1190, block body :1235-1237
        add(3); add(3)
    for _ in range(2):           # models.py:1194-1196 (after the x4 resize)
        add(3); add(5); add(5); add(3)
    add(3, NUMK, 3)              # models.py:1199
    return specs


def init_weights(seed=1234, bias_scale=0.0):
    """glorot_uniform kernels (Keras default), biases zero (or U(+-bias_scale) to exercise bias paths).
    Returns {name: (kernel HWIO float32, bias float32)}."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, k, cin, cout in layer_specs():
        limit = math.sqrt(6.0 / (k * k * cin + k * k * cout))
        w = rng.uniform(-limit, limit, size=(k, k, cin, cout)).astype(np.float32)
        if bias_scale > 0:
            b = rng.uniform(-bias_scale, bias_scale, size=(cout,)).astype(np.float32)
        else:
            b = np.zeros((cout,), dtype=np.float32)
        out[name] = (w, b)
    return out


def n_params(weights=None):
    return sum(k * k * ci * co + co for _, k, ci, co in layer_specs())


def bilinear_x4_tf1(x):
    """tf.image.resize_bilinear legacy (align_corners=False, half_pixel_centers=False), NCHW tensor.
    src = dst * 0.25; lo = floor(src); hi = min(ceil(src), n-1); t = src - lo;
    top = tl + (tr - tl) * tx; bot = bl + (br - bl) * tx; out = top + (bot - top) * ty."""
    n, c, h, w = x.shape

    def axis(n_in):
        dst = torch.arange(4 * n_in, dtype=torch.float64)
        src = dst * 0.25
        lo = torch.floor(src).long()
        hi = torch.clamp(torch.ceil(src).long(), max=n_in - 1)
        t = (src - lo.double()).to(x.dtype)
        return lo, hi, t

    ylo, yhi, ty = axis(h)
    xlo, xhi, tx = axis(w)
    top_rows = x[:, :, ylo, :]
    bot_rows = x[:, :, yhi, :]
    txv = tx.view(1, 1, 1, -1)
    tyv = ty.view(1, 1, -1, 1)
    top = top_rows[:, :, :, xlo] + (top_rows[:, :, :, xhi] - top_rows[:, :, :, xlo]) * txv
    bot = bot_rows[:, :, :, xlo] + (bot_rows[:, :, :, xhi] - bot_rows[:, :, :, xlo]) * txv
    return top + (bot - top) * tyv


class DifvdsrDoubleOracle(torch.nn.Module):
    """Forward graph of models.DifvdsrDouble.create_model (models.py:1159-1222) on NCHW tensors.

    `round_act` optionally rounds activations/weights like the GPU engine does (bf16 operands, fp32
    accumulate) so tests can separate "engine arithmetic" from "kernel bugs"; the parity target itself
    is the plain fp32 (or fp64) run.
    """

    def __init__(self, weights, dtype=torch.float32):
        super().__init__()
        self.names = [s[0] for s in layer_specs()]
        self.ksize = {s[0]: s[1] for s in layer_specs()}
        self.w = torch.nn.ParameterDict()
        self.b = torch.nn.ParameterDict()
        for name in self.names:
            k, b = weights[name]
            # HWIO -> OIHW for F.conv2d (cross-correlation, same as Keras)
            self.w[name] = torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(k)).permute(3, 2, 0, 1).contiguous().to(dtype))
            self.b[name] = torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(b)).to(dtype))

    def conv(self, name, x, relu=False):
        k = self.ksize[name]
        y = F.conv2d(x, self.w[name], self.b[name], padding=(k - 1) // 2)
        return F.relu(y) if relu else y

    def block53(self, x, names):
        # models.py:1248-1270: ini = 0.9*x ; a = conv5(relu(conv3(x))) ; b = conv3(relu(conv5(x)))
        # y = 0.1*(a+b) + ini      (creation order of the four convs: k3, k5, k5, k3)
        ini = 0.9 * x
        a = self.conv(names[1], self.conv(names[0], x, relu=True))
        b = self.conv(names[3], self.conv(names[2], x, relu=True))
        return 0.1 * (a + b) + ini

    def block_light(self, x, names):
        # models.py:1231-1245: y = 0.1*conv3(relu(conv3(x))) + x
        return 0.1 * self.conv(names[1], self.conv(names[0], x, relu=True)) + x

    def forward(self, x_nhwc, return_intermediates=False):
        x = x_nhwc.permute(0, 3, 1, 2)
        inter = {}
        x = self.conv("level1", x, relu=True)
        i = 1
        for blk in range(16):
            x = self.block53(x, self.names[i:i + 4]); i += 4
            if return_intermediates and blk in (0, 15):
                inter["lr53_%d" % blk] = x.permute(0, 2, 3, 1)
        for _ in range(6):
            x = self.block_light(x, self.names[i:i + 2]); i += 2
        if return_intermediates:
            inter["lr_out"] = x.permute(0, 2, 3, 1)
        x = bilinear_x4_tf1(x)
        if return_intermediates:
            inter["hr_in"] = x.permute(0, 2, 3, 1)
        for _ in range(2):
            x = self.block53(x, self.names[i:i + 4]); i += 4
        if return_intermediates:
            inter["hr_out"] = x.permute(0, 2, 3, 1)
        x = self.conv(self.names[i], x, relu=True)
        out = x.permute(0, 2, 3, 1)
        return (out, inter) if return_intermediates else out


def forward_numpy(weights, x_nhwc, dtype=torch.float32, threads=None):
    """model.predict restated: float32 NHWC in [0,1] -> float32 NHWC (models.py:342)."""
    if threads:
        torch.set_num_threads(threads)
    m = DifvdsrDoubleOracle(weights, dtype=dtype)
    with torch.no_grad():
        y = m(torch.from_numpy(np.ascontiguousarray(x_nhwc)).to(dtype))
    return y.to(torch.float32).numpy() if dtype != torch.float64 else y.numpy()


# ------------------------------------------------------------------------------------------------
# training step (models.py:131-157, 1203-1213): mse + Keras-2 Adam
# ------------------------------------------------------------------------------------------------
def mse_loss(pred, target):
    return torch.mean((pred - target) ** 2)


class KerasAdam:
    """keras.optimizers.Adam.get_updates (Keras 2.x): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g^2 ; p = p - lr_t*m/(sqrt(v)+eps)."""

    def __init__(self, params, lr=1e-4, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, epsilon
        self.t = 0
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]

    def step(self, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for p, g, m, v in zip(self.params, grads, self.m, self.v):
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                p.sub_(lr_t * m / (torch.sqrt(v) + self.eps))


def train_step(model, opt, x_nhwc, y_nhwc):
    """One train_on_batch: forward, mse, backward, Adam.  Returns (loss, grads dict in HWIO/bias layout)."""
    params = list(model.parameters())
    pred = model(x_nhwc)
    loss = mse_loss(pred, y_nhwc)
    grads = torch.autograd.grad(loss, params)
    opt.step(grads)
    return float(loss), grads
