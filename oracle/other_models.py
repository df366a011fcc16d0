"""CPU restatement of the reference's two older graphs: Difvdsr4 (models.py:992-1142) and Difvdsr
(models.py:1274-1357).

TEST INFRASTRUCTURE (see oracle/__init__.py): only tests/ may import this.

PARITY UNPINNED, like oracle/model.py: Keras 2 / TensorFlow 1 are not installable here, the reference ships neither
weights (`weights_Difvdsr2scale/...`, `weights_Difvdsr/weights-23-0.96.h5` are external) nor outputs for these
models, and `Difvdsr.create_model` forces `load_weights=True` (models.py:1322) so the reference cannot even build it
without that file.  The restatement follows the source line by line:

Difvdsr4 (numk = 256, models.py:1017):
    level1   Convolution2D(256,(1,1),relu)                                     :1024
    6  x _residual_block_light0   y = 0.1*conv3(LeakyReLU(0.001)(conv3(x))) + x   :1030-1032, body :1127-1142
    Lambda(resize2bil)  tf.image.resize_bilinear x2 (TF1 legacy sampling)      :1034, :932-940
    xInp = x ; 20 x _residual_block_light  y = 0.1*conv3(relu(conv3(x))) + x   :1035-1038, body :1110-1125
    x = Add([x, xInp])                                                         :1039
    Lambda(resize2bil)                                                         :1041
    6  x _residual_block_light                                                 :1042-1044
    Conv2D(3,(3,3),relu)                                                       :1047
Difvdsr (numk = 192, same resolution, models.py:1297):
    level1   Convolution2D(192,(3,3),relu, trainable=False)                    :1304
    32 x _residual_block                                                       :1305-1306, body :1336-1357
        a = conv3(relu(conv3(x))) ; d = a - x ; e = conv3(LeakyReLU(0.2)(conv3(d)))
        y = 0.1*(d + e + a) + x             (resizeRes01 = scalar_mul(0.1), models.py:1432-1434)
    Conv2D(3,(3,3),relu)                                                       :1308
Keras auto-names in creation order: 'level1', 'conv2d_1', ...; LeakyReLU / Activation / Lambda / Add / Subtract
layers have no weights.  keras.layers.LeakyReLU(alpha): f(x) = x if x >= 0 else alpha*x [lib].
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def difvdsr4_specs(numk=256):
    specs = [("level1", 1, 3, numk)]
    for i in range(1, 2 * (6 + 20 + 6) + 1):
        specs.append(("conv2d_%d" % i, 3, numk, numk))
    specs.append(("conv2d_%d" % (len(specs)), 3, numk, 3))
    return specs


def difvdsr_specs(numk=192):
    specs = [("level1", 3, 3, numk)]
    for i in range(1, 4 * 32 + 1):
        specs.append(("conv2d_%d" % i, 3, numk, numk))
    specs.append(("conv2d_%d" % (len(specs)), 3, numk, 3))
    return specs


def init_weights(specs, seed=1234, bias_scale=0.0, gain=1.0):
    """glorot_uniform kernels x gain, zero (or U(+-bias_scale)) biases: {name: (HWIO float32, bias float32)}."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, k, cin, cout in specs:
        limit = gain * math.sqrt(6.0 / (k * k * cin + k * k * cout))
        w = rng.uniform(-limit, limit, size=(k, k, cin, cout)).astype(np.float32)
        b = (rng.uniform(-bias_scale, bias_scale, size=(cout,)) if bias_scale > 0 else np.zeros((cout,))).astype(np.float32)
        out[name] = (w, b)
    return out


def bilinear_tf1(x, r):
    """tf.image.resize_bilinear(x, [r*h, r*w]) legacy sampling (align_corners=False, no half-pixel centres), NCHW:
    src = dst / r ; lo = floor(src) ; hi = min(lo + 1, n - 1) ; t = src - lo ;
    top = tl + (tr - tl)*tx ; bot = bl + (br - bl)*tx ; out = top + (bot - top)*ty."""
    n, c, h, w = x.shape

    def axis(n_in):
        dst = torch.arange(r * n_in, dtype=torch.float64)
        src = dst / r
        lo = torch.floor(src).long()
        hi = torch.clamp(lo + 1, max=n_in - 1)
        return lo, hi, (src - lo.double()).to(x.dtype)

    ylo, yhi, ty = axis(h)
    xlo, xhi, tx = axis(w)
    top_rows, bot_rows = x[:, :, ylo, :], x[:, :, yhi, :]
    txv, tyv = tx.view(1, 1, 1, -1), ty.view(1, 1, -1, 1)
    top = top_rows[:, :, :, xlo] + (top_rows[:, :, :, xhi] - top_rows[:, :, :, xlo]) * txv
    bot = bot_rows[:, :, :, xlo] + (bot_rows[:, :, :, xhi] - bot_rows[:, :, :, xlo]) * txv
    return top + (bot - top) * tyv


class _Net:
    def __init__(self, specs, weights, dtype):
        self.names = [s[0] for s in specs]
        self.k = {s[0]: s[1] for s in specs}
        self.w = {n: torch.from_numpy(np.ascontiguousarray(weights[n][0])).permute(3, 2, 0, 1).contiguous().to(dtype)
                  for n in self.names}
        self.b = {n: torch.from_numpy(np.ascontiguousarray(weights[n][1])).to(dtype) for n in self.names}
        self.dtype = dtype

    def conv(self, name, x):
        return F.conv2d(x, self.w[name], self.b[name], padding=(self.k[name] - 1) // 2)


def _graph_difvdsr4(net, x):
    x = F.relu(net.conv("level1", x))
    i = 1
    for _ in range(6):
        t = F.leaky_relu(net.conv(net.names[i], x), 0.001)
        x = 0.1 * net.conv(net.names[i + 1], t) + x
        i += 2
    x = bilinear_tf1(x, 2)
    x_inp = x
    for _ in range(20):
        x = 0.1 * net.conv(net.names[i + 1], F.relu(net.conv(net.names[i], x))) + x
        i += 2
    x = x + x_inp
    x = bilinear_tf1(x, 2)
    for _ in range(6):
        x = 0.1 * net.conv(net.names[i + 1], F.relu(net.conv(net.names[i], x))) + x
        i += 2
    return F.relu(net.conv(net.names[i], x))


def _graph_difvdsr(net, x):
    x = F.relu(net.conv("level1", x))
    i = 1
    for _ in range(32):
        a = net.conv(net.names[i + 1], F.relu(net.conv(net.names[i], x)))
        d = a - x
        e = net.conv(net.names[i + 3], F.leaky_relu(net.conv(net.names[i + 2], d), 0.2))
        x = 0.1 * (d + e + a) + x
        i += 4
    return F.relu(net.conv(net.names[i], x))


def forward_difvdsr4(weights, x_nhwc, dtype=torch.float32, numk=256):
    net = _Net(difvdsr4_specs(numk), weights, dtype)
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(x_nhwc)).to(dtype).permute(0, 3, 1, 2)
        return _graph_difvdsr4(net, x).permute(0, 2, 3, 1).to(torch.float32).numpy()


def forward_difvdsr(weights, x_nhwc, dtype=torch.float32, numk=192):
    net = _Net(difvdsr_specs(numk), weights, dtype)
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(x_nhwc)).to(dtype).permute(0, 3, 1, 2)
        return _graph_difvdsr(net, x).permute(0, 2, 3, 1).to(torch.float32).numpy()


def loss_and_grads(arch, weights, x_nhwc, y_nhwc, dtype=torch.float64):
    """compile(loss='mse') + one backward pass (models.py:1057-1058, 1318-1319): (loss, {name: (d loss / d kernel in
    HWIO, d loss / d bias)}) by torch autograd on the same graph.  Difvdsr's 'level1' is trainable=False in the
    reference (models.py:1304): its entry is zeros."""
    specs = difvdsr4_specs() if arch == "difvdsr4" else difvdsr_specs()
    net = _Net(specs, weights, dtype)
    frozen = {"level1"} if arch == "difvdsr" else set()
    params = []
    for n in net.names:
        if n not in frozen:
            net.w[n].requires_grad_(True)
            net.b[n].requires_grad_(True)
            params += [net.w[n], net.b[n]]
    x = torch.from_numpy(np.ascontiguousarray(x_nhwc)).to(dtype).permute(0, 3, 1, 2)
    y = torch.from_numpy(np.ascontiguousarray(y_nhwc)).to(dtype).permute(0, 3, 1, 2)
    pred = (_graph_difvdsr4 if arch == "difvdsr4" else _graph_difvdsr)(net, x)
    loss = torch.mean((pred - y) ** 2)
    grads = torch.autograd.grad(loss, params)
    out, it = {}, iter(grads)
    for n in net.names:
        if n in frozen:
            out[n] = (np.zeros(tuple(net.w[n].permute(2, 3, 1, 0).shape), np.float32),
                      np.zeros(tuple(net.b[n].shape), np.float32))
        else:
            gw, gb = next(it), next(it)
            out[n] = (gw.permute(2, 3, 1, 0).contiguous().to(torch.float32).numpy(), gb.to(torch.float32).numpy())
    return float(loss.detach()), out, pred.detach().permute(0, 2, 3, 1).to(torch.float32).numpy()
