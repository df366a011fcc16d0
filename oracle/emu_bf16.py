"""The DifvdsrDouble oracle graph with the GPU engine's ROUNDING POINTS (bf16 operands, fp32 accumulation).

TEST INFRASTRUCTURE (see oracle/model.py).  This is not a second parity target -- the target stays the plain fp32
graph of oracle/model.py (reference models.py:1159-1270, loss / optimizer :1203-1213).  It exists to make the
training-gradient check sharp: compared with the fp32 graph, bf16 operands cost a few 1e-2 of relative error per
layer, which would hide a transposed tap or a dropped halo column; compared with THIS graph (same operands, same
rounding of saved activations and stored gradients) the engine's gradients must agree to ~1e-3, the remainder being
fp32 summation order and double rounding.

Rounding points mirrored (sr100/train.py, csrc/conv_tc.cu epilogues):
  forward   conv operands (activations and packed weights) are bf16; a 128-wide epilogue rounds the fp32 accumulator
            to bf16 before bias / alpha / residual (`acc_round`); t1 / t2 are stored bf16; the LR residual stream is
            fp32 with a bf16 operand copy, the HR stream (training graph) is bf16 only; the bilinear writes bf16;
            the 3-channel tail keeps its fp32 accumulator.
  backward  every stored gradient tensor (g, a1, a2, the tail's im2col'ed loss gradient) is bf16; the LR stream
            gradient additionally lives in fp32 (g32) and only its operand copy is rounded; the 0.1 of a block tail is
            applied after the bf16 gradient operand (dgrad alpha / wgrad scale).
Autograd carries this through three identities: rb_fb (round forward and the incoming gradient), rb_f (forward
only: weights, already-rounded tensors) and rgrad (identity forward, bf16-round the gradient)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .model import bilinear_x4_tf1, layer_specs


def _bf16(t):
    return t.to(torch.bfloat16).to(t.dtype)


class _RoundFwdBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


rb_fb, rb_f, rgrad = _RoundFwdBwd.apply, _RoundFwd.apply, _RoundGrad.apply


class DifvdsrDoubleBf16Emu(torch.nn.Module):
    """Same parameters / names as oracle.model.DifvdsrDoubleOracle; float64 arithmetic between the rounding points."""

    def __init__(self, weights, acc_round=True, hr_stream_fp32=False):
        super().__init__()
        self.names = [s[0] for s in layer_specs()]
        self.ksize = {s[0]: s[1] for s in layer_specs()}
        self.acc_round, self.hr_stream_fp32 = acc_round, hr_stream_fp32
        self.w = torch.nn.ParameterDict()
        self.b = torch.nn.ParameterDict()
        for name in self.names:
            k, b = weights[name]
            self.w[name] = torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(k)).permute(3, 2, 0, 1).contiguous().double())
            self.b[name] = torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(b)).double())

    def acc(self, name, x_op):
        """fp32-accumulator stand-in: conv of a bf16 operand tensor with the bf16 weights, no bias."""
        k = self.ksize[name]
        return F.conv2d(x_op, rb_f(self.w[name]), None, padding=(k - 1) // 2)

    def racc(self, a):
        return rb_f(a) if self.acc_round else a      # forward-only: the backward of the epilogue's rounding is identity

    def branch(self, name, x_op):
        """conv + bias + ReLU -> stored bf16 (t1 / t2); its gradient (a1 / a2) is stored bf16 too."""
        return rb_fb(F.relu(self.racc(self.acc(name, x_op)) + self.b[name].view(1, -1, 1, 1)))

    def block(self, s, names, stream_fp32, beta):
        """s: the residual stream (fp32 values at LR, bf16 values at HR).  Returns the new stream."""
        x_op = rb_fb(s)        # operand copy: rounds forward; the summed operand-path gradient is rounded once (bf16(acc))
        if len(names) == 4:    # 5/3 block: creation order k3, k5, k5, k3 (models.py:1253-1259)
            z = self.acc(names[1], self.branch(names[0], x_op)) + self.acc(names[3], self.branch(names[2], x_op))
            bias = (self.b[names[1]] + self.b[names[3]]).view(1, -1, 1, 1)
        else:                  # light block (models.py:1235-1237)
            z = self.acc(names[1], self.branch(names[0], x_op))
            bias = self.b[names[1]].view(1, -1, 1, 1)
        # 0.1 * (z + bias): the gradient operand of the tail convs is the bf16 copy of the stream gradient
        upd = rgrad(0.1 * (self.racc(z) + bias))
        if stream_fp32:
            return upd + beta * s                  # fp32 stream; residual-path gradient stays fp32 (g32)
        return rb_fb(upd + beta * x_op_value(s))   # bf16 stream: value and gradient rounded every block

    def forward(self, x_nhwc):
        x = x_nhwc.permute(0, 3, 1, 2).double()
        # head: CUDA cores, fp32 weights and input, fp32 stream (+ bf16 operand copy taken by the first block)
        s = F.relu(F.conv2d(x, self.w["level1"], self.b["level1"]))
        i = 1
        for _ in range(16):
            s = self.block(s, self.names[i:i + 4], True, 0.9); i += 4
        for _ in range(6):
            s = self.block(s, self.names[i:i + 2], True, 1.0); i += 2
        s = bilinear_x4_tf1(s)
        if not self.hr_stream_fp32:
            s = rb_f(s)            # the bilinear writes bf16; its gradient (gsh32) is fp32
        for _ in range(2):
            s = self.block(s, self.names[i:i + 4], self.hr_stream_fp32, 0.9); i += 4
        # tail: 3 output channels, fp32 accumulator kept; the loss gradient at the pre-activation is stored bf16
        k = self.ksize[self.names[i]]
        pre = rgrad(F.conv2d(rb_fb(s), rb_f(self.w[self.names[i]]), None, padding=(k - 1) // 2)
                    + self.b[self.names[i]].view(1, -1, 1, 1))
        return F.relu(pre).permute(0, 2, 3, 1)


def x_op_value(s):
    """At HR the residual operand of the block tail is the bf16 stream itself (already bf16 values): identity."""
    return s


def gradients(weights, x_nhwc, y_nhwc, **kw):
    """(loss, {name: (dW HWIO, db)}) of mse(model(x), y) with the engine's rounding points."""
    m = DifvdsrDoubleBf16Emu(weights, **kw)
    pred = m(torch.from_numpy(np.ascontiguousarray(x_nhwc)))
    loss = torch.mean((pred - torch.from_numpy(np.ascontiguousarray(y_nhwc)).double()) ** 2)
    params = list(m.parameters())
    grads = torch.autograd.grad(loss, params)
    out = {}
    for (pname, _), g in zip(m.named_parameters(), grads):
        kind, lname = pname.split(".")
        out.setdefault(lname, {})[kind] = g.numpy()
    return float(loss.detach()), {n: (np.transpose(out[n]["w"], (2, 3, 1, 0)), out[n]["b"]) for n in m.names}
