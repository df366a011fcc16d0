"""Training step of the reference's two older graphs on sm_100a: Difvdsr4 (256 channels, models.py:992-1082) and
Difvdsr (192 channels, models.py:1274-1357) -- `fit` of both classes is `BaseSuperResolutionModel.fit`
(models.py:131-157, :1079-1080, :1332-1333) on a graph compiled with loss='mse', Adam(1e-4, 0.9) (models.py:1057-1058,
:1318-1319).

Same kernels as the DifvdsrDouble trainer (sr100.train): the input gradient of a conv is the forward tensor-core
kernel with 180-degree-rotated, cin <-> cout-swapped weights, the filter gradient is csrc/wgrad_tc.cu, bias gradients
are column sums.  A C-channel tensor is two 128-channel planes (sr100.planenet), so a C -> C layer costs two
two-source launches per direction and four 128 x 128 filter-gradient launches, each writing its dense plane block of
the gradient arena (same layout as PlaneNet.param_arena: one Adam launch updates everything, padded entries have zero
gradients and stay zero).

    light / light0 block   t = act(conv_a(x)) ; y = 0.1 * conv_b(t) + x            act = ReLU / LeakyReLU(0.001)
        g_t = 0.1 * dgrad_b(g_y) * act'(t)          g_x = dgrad_a(g_t) + g_y
    difference block       t = relu(conv1(x)) ; d = conv2(t) - x ; u = LeakyReLU(0.2)(conv3(d)) ;
                           y = 0.1 * conv4(u) + 0.2 * d + 1.1 * x                  (= 0.1 * (d + e + a) + x, :1352-1355)
        g_u = 0.1 * dgrad4(g_y) * lrelu'(u)          g_d = dgrad3(g_u) + 0.2 * g_y
        g_t = dgrad2(g_d) * relu'(t)                 g_x = dgrad1(g_t) + 1.1 * g_y - g_d
LeakyReLU keeps the sign of its input, so its saved OUTPUT is the mask (sr_conv_desc.relu_mask_slope).  The tail conv
(C -> 3) is the K = 27 im2col formulation of sr100.train per input plane; bilinear x2 has its adjoint
(sr_bilinear2_bwd); Difvdsr's 3-channel head is `trainable=False` in the reference (models.py:1304) and gets no
gradient.  Operands bf16, accumulation fp32, master weights / gradients / Adam state fp32."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .engine import _Plan
from .planenet import NUMK
from .train import _WgradPlan


class _PlaneTrainGraph:
    """Buffers and launch lists (forward with saved activations, backward) of one (NB, H, W) minibatch shape."""

    def __init__(self, tr, NB, H, W):
        self.tr, self.net = tr, tr.net
        net = tr.net
        self.lib, self.dev = net.lib, net.device
        self.NB, self.H, self.W = NB, H, W
        s = net.scale
        self.x_in = torch.empty(NB, H, W, 3, device=self.dev, dtype=torch.float32)
        self.y_true = torch.empty(NB, s * H, s * W, 3, device=self.dev, dtype=torch.float32)
        self.out = torch.empty(NB, s * H, s * W, 3, device=self.dev, dtype=torch.float32)
        self.loss_sum = torch.zeros(1, device=self.dev, dtype=torch.float64)
        self.n_local = self.out.numel()
        self.fwd, self.bwd, self.keep = [], [], []
        self.flops = 0.0
        self.cuda_graph, self.ran_eager = None, False
        (self._difvdsr4 if net.arch == "difvdsr4" else self._difvdsr)()

    # ------------------------------------------------------------------ building blocks
    def planes(self, shape, dtype=torch.bfloat16, zero=False):
        mk = torch.zeros if zero else torch.empty
        t = [mk(*shape, NUMK, device=self.dev, dtype=dtype) for _ in range(2)]
        self.keep.append(t)
        return t

    def conv2(self, dst, name, src16, shape, flip=False, out16=None, out32=None, relu=0, slope=0.0, alpha=1.0,
              beta=0.0, res32=None, mask=None, mask_slope=0.0):
        """A C -> C conv (flip: its input gradient) as two two-source launches, one per output plane."""
        net, tr = self.net, self.tr
        for o in range(2):
            d = L.ConvDesc()
            d.nsrc = 2
            for s_ in range(2):
                d.in_[s_] = src16[s_].data_ptr()
                d.wpacked[s_] = (tr.packed_t if flip else net.packed)[(name, s_, o)].data_ptr()
                d.ksize[s_] = net.ksize[name]
            d.NB, d.H, d.W = shape
            d.cin, d.cout = NUMK, NUMK
            d.cin_valid[1] = net.C - NUMK if net.skip_zero_k else 0
            d.bias = None if flip else net.bias[(name, o)].data_ptr()
            d.alpha, d.beta, d.relu, d.leaky_slope = alpha, beta, relu, slope
            d.res_f32 = res32[o].data_ptr() if res32 is not None else None
            d.out_bf16 = out16[o].data_ptr() if out16 is not None else None
            d.out_f32 = out32[o].data_ptr() if out32 is not None else None
            if mask is not None:
                d.relu_mask_bf16, d.relu_mask_slope = mask[o].data_ptr(), mask_slope
            d.a_mode, d.nacc, d.pair = 0, 2, 1
            p = _Plan(net.lib, d)
            self.flops += p.flops
            dst.append(p.run)

    def wgrad2(self, name, x16, g16, shape, scale):
        tr = self.tr
        for i in range(2):
            for j in range(2):
                p = _WgradPlan(self.lib, x16[i], g16[j], shape, self.net.ksize[name], scale, tr.gblock(name, i, j),
                               tr.workspace)
                self.flops += p.flops
                self.bwd.append(p.run)

    def colsum2(self, name, g16, shape, scale):
        npix = shape[0] * shape[1] * shape[2]
        lib = self.lib
        for j in range(2):
            db = self.tr.gbias(name, j)
            self.bwd.append(lambda st, j=j, db=db: L.check(lib.sr_colsum_bf16(L.ptr(g16[j]), npix, scale, L.ptr(db), st)))

    def axpby(self, dst, x32, y32, a, b, out32=None, out16=None):
        lib = self.lib
        for j in range(2):
            n = x32[j].numel()
            dst.append(lambda st, j=j, n=n: L.check(lib.sr_axpby_f32(
                L.ptr(x32[j]), L.ptr(y32[j]), a, b, n, L.ptr(out32[j]) if out32 is not None else None,
                L.ptr(out16[j]) if out16 is not None else None, st)))

    def upsample2(self, s32, shape):
        NB, H, W = shape
        up = (NB, 2 * H, 2 * W)
        o16, o32 = self.planes(up), self.planes(up, torch.float32)
        lib = self.lib
        for j in range(2):
            self.fwd.append(lambda st, j=j: L.check(lib.sr_bilinear2_fwd(
                L.ptr(s32[j]), 0, NB, H, W, NUMK, L.ptr(o16[j]), L.ptr(o32[j]), st)))
        return o16, o32, up

    def upsample2_bwd(self, g32_hi, shape_lo, g32_lo, g16_lo):
        NB, H, W = shape_lo
        lib = self.lib
        for j in range(2):
            n = NB * H * W * NUMK
            self.bwd.append(lambda st, j=j: L.check(lib.sr_bilinear2_bwd(L.ptr(g32_hi[j]), NB, H, W, NUMK,
                                                                        L.ptr(g32_lo[j]), st)))
            self.bwd.append(lambda st, j=j, n=n: L.check(lib.sr_cast_f32_to_bf16(L.ptr(g32_lo[j]), n, L.ptr(g16_lo[j]), st)))

    # ------------------------------------------------------------------ the tail conv (C -> 3, ReLU) and the loss
    def tail_fwd(self, src16, shape):
        net = self.net
        d = L.ConvDesc()
        d.nsrc = 2
        for i in range(2):
            d.in_[i] = src16[i].data_ptr()
            d.wpacked[i] = net.packed[(net.tail, i, 0)].data_ptr()
            d.ksize[i] = 3
        d.NB, d.H, d.W = shape
        d.cin, d.cout = NUMK, 3
        d.cin_valid[1] = net.C - NUMK if net.skip_zero_k else 0
        d.bias = net.bias[(net.tail, 0)].data_ptr()
        d.alpha, d.beta, d.relu = 1.0, 0.0, 1
        d.out_f32 = self.out.data_ptr()
        d.a_mode, d.nacc, d.pair = 0, 2, 0
        p = _Plan(net.lib, d)
        self.flops += p.flops
        self.fwd.append(p.run)

    def tail_bwd(self, src16, shape, g16, g32):
        """loss = mse(out, y): the loss gradient through the tail's ReLU is written as the im2col of the tail's
        backward (27 of 128 bf16 channels), which makes both tail gradients 1x1 problems per input plane."""
        net, tr, lib = self.net, self.tr, self.lib
        NB, HH, WW = shape
        gcol = torch.zeros(NB, HH, WW, NUMK, device=self.dev, dtype=torch.bfloat16)
        self.keep.append(gcol)
        tail_b = tr.gbias(net.tail, 0)
        self.bwd.append(lambda st: L.check(lib.sr_mse_tail_grad_col(
            L.ptr(self.out), L.ptr(self.y_true), NB, HH, WW, self.n_local, L.ptr(gcol), L.ptr(self.loss_sum),
            L.ptr(tail_b), st)))
        for i in range(2):
            d128 = tr.tail_d128[i]
            p = _WgradPlan(lib, src16[i], gcol, shape, 1, 1.0, d128, tr.workspace)          # D[ci][j], j = (ky,kx,co)
            self.flops += 2.0 * NB * HH * WW * 27 * NUMK
            self.bwd.append(p.run)
            gw = tr.gblock(net.tail, i, 0)                                                  # [3,3,128,3]
            self.bwd.append(lambda st, gw=gw, d128=d128: gw.copy_(
                d128.view(NUMK, NUMK)[:, :27].reshape(NUMK, 3, 3, 3).permute(1, 2, 0, 3)))
            d = L.ConvDesc()                                                                # dgrad: 1x1 conv, B[j][ci]
            d.nsrc = 1
            d.in_[0], d.wpacked[0], d.ksize[0] = gcol.data_ptr(), tr.tail_colw_packed[i].data_ptr(), 1
            d.NB, d.H, d.W, d.cin, d.cout = NB, HH, WW, NUMK, NUMK
            d.alpha, d.beta, d.relu = 1.0, 0.0, 0
            d.out_bf16, d.out_f32 = g16[i].data_ptr(), g32[i].data_ptr()
            d.a_mode, d.nacc, d.pair = 0, 2, 1
            pl = _Plan(lib, d)
            self.flops += 2.0 * NB * HH * WW * 27 * NUMK
            self.bwd.append(pl.run)

    # ------------------------------------------------------------------ light blocks (Difvdsr4)
    def light_fwd(self, i, x16, s32, shape, leaky=None):
        names = self.net.names
        t16, y16 = self.planes(shape), self.planes(shape)
        if leaky is None:
            self.conv2(self.fwd, names[i], x16, shape, out16=t16, relu=1)
        else:
            self.conv2(self.fwd, names[i], x16, shape, out16=t16, relu=2, slope=leaky)
        self.conv2(self.fwd, names[i + 1], t16, shape, out16=y16, out32=s32, alpha=0.1, beta=1.0, res32=s32)
        return t16, y16

    def light_bwd(self, i, x16, t16, shape, g16, g32, gt16, leaky=None):
        names = self.net.names
        na, nb = names[i], names[i + 1]
        self.conv2(self.bwd, nb, g16, shape, flip=True, out16=gt16, alpha=0.1, mask=t16,
                   mask_slope=0.0 if leaky is None else leaky)
        self.wgrad2(nb, t16, g16, shape, 0.1)
        self.colsum2(nb, g16, shape, 0.1)
        self.conv2(self.bwd, na, gt16, shape, flip=True, out16=g16, out32=g32, alpha=1.0, beta=1.0, res32=g32)
        self.wgrad2(na, x16, gt16, shape, 1.0)
        self.colsum2(na, gt16, shape, 1.0)

    def _difvdsr4(self):
        net, lib, tr = self.net, self.lib, self.tr
        NB, H, W = self.NB, self.H, self.W
        f32 = torch.float32
        lr = (NB, H, W)
        npix = NB * H * W
        x16, s32 = self.planes(lr), self.planes(lr, f32)
        head16 = x16
        for j in range(2):                                     # level1: 1x1, 3 -> 256, ReLU (models.py:1024)
            self.fwd.append(lambda st, j=j, o16=x16[j], o32=s32[j]: L.check(lib.sr_head1x1_fwd(
                L.ptr(self.x_in), L.ptr(net.head_w[j]), L.ptr(net.bias[("level1", j)]), npix, L.ptr(o16), L.ptr(o32), st)))
        stages = []                                            # (first layer index, x16, t16, shape, leaky)
        i, shape = 1, lr
        for _ in range(6):                                     # :1030-1032
            t16, y16 = self.light_fwd(i, x16, s32, shape, leaky=0.001)
            stages.append((i, x16, t16, shape, 0.001))
            x16, i = y16, i + 2
        n_lr = len(stages)
        x16, s32, shape = self.upsample2(s32, shape)           # :1034
        mid = shape
        xinp = self.planes(shape, f32)                         # xInp = x (:1035)
        for j in range(2):
            self.fwd.append(lambda st, j=j, a=xinp, b=s32: a[j].copy_(b[j]))
        for _ in range(20):                                    # :1036-1038
            t16, y16 = self.light_fwd(i, x16, s32, shape)
            stages.append((i, x16, t16, shape, None))
            x16, i = y16, i + 2
        n_mid = len(stages)
        self.axpby(self.fwd, s32, xinp, 1.0, 1.0, s32)         # Add([x, xInp]) (:1039); only its x2 upsampling is consumed
        x16, s32, shape = self.upsample2(s32, shape)           # :1041
        hr = shape
        for _ in range(6):                                     # :1042-1044
            t16, y16 = self.light_fwd(i, x16, s32, shape)
            stages.append((i, x16, t16, shape, None))
            x16, i = y16, i + 2
        self.tail_fwd(x16, shape)                              # :1047

        # ---------------------------------------------------------------- backward
        g16h, g32h, gt16h = self.planes(hr), self.planes(hr, f32), self.planes(hr)
        self.tail_bwd(x16, hr, g16h, g32h)
        for st_ in reversed(stages[n_mid:]):
            self.light_bwd(st_[0], st_[1], st_[2], st_[3], g16h, g32h, gt16h, st_[4])
        g16m, g32m, gt16m = self.planes(mid), self.planes(mid, f32), self.planes(mid)
        self.upsample2_bwd(g32h, mid, g32m, g16m)
        gx32 = self.planes(mid, f32)                           # the gradient that reaches xInp directly
        for j in range(2):
            self.bwd.append(lambda st, j=j: gx32[j].copy_(g32m[j]))
        for st_ in reversed(stages[n_lr:n_mid]):
            self.light_bwd(st_[0], st_[1], st_[2], st_[3], g16m, g32m, gt16m, st_[4])
        self.axpby(self.bwd, g32m, gx32, 1.0, 1.0, g32m)
        g16l, g32l, gt16l = self.planes(lr), self.planes(lr, f32), self.planes(lr)
        self.upsample2_bwd(g32m, lr, g32l, g16l)
        for st_ in reversed(stages[:n_lr]):
            self.light_bwd(st_[0], st_[1], st_[2], st_[3], g16l, g32l, gt16l, st_[4])
        for j in range(2):                                     # head: g0 = g * (act > 0); dw = x^T g0; db = colsum g0
            hw, hb = tr.gblock("level1", 0, j), tr.gbias("level1", j)
            self.bwd.append(lambda st, j=j, hw=hw, hb=hb: L.check(lib.sr_head1x1_bwd(
                L.ptr(self.x_in), L.ptr(head16[j]), L.ptr(g32l[j]), None, npix, L.ptr(hw), L.ptr(hb), st)))

    # ------------------------------------------------------------------ difference blocks (Difvdsr)
    def _difvdsr(self):
        net, lib = self.net, self.lib
        NB, H, W = self.NB, self.H, self.W
        f32 = torch.float32
        shape = (NB, H, W)
        Cc = net.C
        x16, s32 = self.planes(shape, zero=True), self.planes(shape, f32, zero=True)
        d32, r32 = self.planes(shape, f32), self.planes(shape, f32)
        head = torch.empty(NB, H, W, Cc, device=self.dev, dtype=f32)
        self.keep.append(head)
        w0, b0 = net.blocks[("level1", 0, 0)], net.bias[("level1", 0)]
        self.fwd.append(lambda st: L.check(lib.sr_conv2d_direct(          # level1: 3x3, 3 -> 192, ReLU (:1304)
            L.ptr(self.x_in), 0, L.ptr(w0), 0, L.ptr(b0), NB, H, W, 3, Cc, 3, 1, 1, 0, 0, L.ptr(head), st)))
        x0 = x16

        def split(st):
            s32[0].copy_(head[..., :NUMK])
            s32[1][..., :Cc - NUMK].copy_(head[..., NUMK:])
            x0[0].copy_(s32[0])
            x0[1].copy_(s32[1])
        self.fwd.append(split)
        names = net.names
        stages = []
        i = 1
        for _ in range(32):                                                # :1305-1306, body :1336-1357
            t16, d16, u16, y16 = self.planes(shape), self.planes(shape), self.planes(shape), self.planes(shape)
            self.conv2(self.fwd, names[i], x16, shape, out16=t16, relu=1)
            self.conv2(self.fwd, names[i + 1], t16, shape, out16=d16, out32=d32, alpha=1.0, beta=-1.0, res32=s32)
            self.axpby(self.fwd, s32, d32, 1.1, 0.2, r32)
            self.conv2(self.fwd, names[i + 2], d16, shape, out16=u16, relu=2, slope=0.2)
            self.conv2(self.fwd, names[i + 3], u16, shape, out16=y16, out32=s32, alpha=0.1, beta=1.0, res32=r32)
            stages.append((i, x16, t16, d16, u16))
            x16, i = y16, i + 4
        self.tail_fwd(x16, shape)                                          # :1308

        # ---------------------------------------------------------------- backward
        g16, g32 = self.planes(shape), self.planes(shape, f32)
        gu16, gd16, gd32 = self.planes(shape), self.planes(shape), self.planes(shape, f32)
        gt16, r2 = self.planes(shape), self.planes(shape, f32)
        self.tail_bwd(x16, shape, g16, g32)
        for (i, xb, tb, db, ub) in reversed(stages):
            n1, n2, n3, n4 = names[i], names[i + 1], names[i + 2], names[i + 3]
            self.conv2(self.bwd, n4, g16, shape, flip=True, out16=gu16, alpha=0.1, mask=ub, mask_slope=0.2)
            self.wgrad2(n4, ub, g16, shape, 0.1)
            self.colsum2(n4, g16, shape, 0.1)
            self.conv2(self.bwd, n3, gu16, shape, flip=True, out16=gd16, out32=gd32, alpha=1.0, beta=0.2, res32=g32)
            self.wgrad2(n3, db, gu16, shape, 1.0)
            self.colsum2(n3, gu16, shape, 1.0)
            self.conv2(self.bwd, n2, gd16, shape, flip=True, out16=gt16, alpha=1.0, mask=tb)
            self.wgrad2(n2, tb, gd16, shape, 1.0)
            self.colsum2(n2, gd16, shape, 1.0)
            self.axpby(self.bwd, g32, gd32, 1.1, -1.0, r2)
            self.conv2(self.bwd, n1, gt16, shape, flip=True, out16=g16, out32=g32, alpha=1.0, beta=1.0, res32=r2)
            self.wgrad2(n1, xb, gt16, shape, 1.0)
            self.colsum2(n1, gt16, shape, 1.0)
        # level1 is trainable=False in the reference (models.py:1304): no head gradient


class PlaneTrainer:
    """train_on_batch / evaluate for a kmodel.Model over a PlaneNet (Keras Model.train_on_batch: returns the loss)."""

    def __init__(self, net, lr=1e-4, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.net, self.lib = net, net.lib
        self.lr, self.beta_1, self.beta_2, self.epsilon = float(lr), float(beta_1), float(beta_2), float(epsilon)
        dev, n = net.device, net.n_arena
        self.grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.t = 0
        self.workspace = torch.empty(self.lib.sr_wgrad_workspace_bytes(), dtype=torch.uint8, device=dev)
        self._gviews = {}
        for name, pieces in net.layout.items():
            for kind, key, o, shape in pieces:
                view = self.grads[o:o + int(np.prod(shape))].view(*shape)
                self._gviews[(name, kind) + (key if isinstance(key, tuple) else (key,))] = view
        # input-gradient weights: packed_t[(name, j, i)] = plane block (i, j) rotated / transposed
        self.packed_t = {}
        for (name, i, j), blk in net.blocks.items():
            if blk.dim() == 4 and blk.shape[2] == NUMK and blk.shape[3] == NUMK:
                self.packed_t[(name, j, i)] = torch.empty(self.lib.sr_packed_weight_bytes(net.ksize[name], NUMK),
                                                          dtype=torch.uint8, device=dev)
        self._pack_table_t = net.make_pack_table(self.packed_t, flip=True)
        self.tail_d128 = [torch.zeros(1, 1, NUMK, NUMK, dtype=torch.float32, device=dev) for _ in range(2)]
        self.tail_colw = [torch.zeros(1, 1, NUMK, NUMK, dtype=torch.float32, device=dev) for _ in range(2)]
        self.tail_colw_packed = [torch.empty(self.lib.sr_packed_weight_bytes(1, NUMK), dtype=torch.uint8, device=dev)
                                 for _ in range(2)]
        self._graphs = {}
        self.max_graphs = 2
        self.repack_t()

    def gblock(self, name, i, j):
        return self._gviews[(name, "w", i, j)]

    def gbias(self, name, j):
        return self._gviews[(name, "b", j)]

    def grads_dict(self):
        """{name: (kernel gradient HWIO, bias gradient)} assembled from the plane blocks (numpy)."""
        net, Cc = self.net, self.net.C
        out = {}
        for name, k, cin, cout in net.specs:
            if cin == Cc and cout == Cc:
                w = torch.zeros(k, k, 2 * NUMK, 2 * NUMK, device=net.device)
                for i in range(2):
                    for j in range(2):
                        w[:, :, i * NUMK:(i + 1) * NUMK, j * NUMK:(j + 1) * NUMK] = self.gblock(name, i, j)
                b = torch.cat([self.gbias(name, 0), self.gbias(name, 1)])
                out[name] = (w[:, :, :Cc, :Cc].cpu().numpy(), b[:Cc].cpu().numpy())
            elif cin == Cc:
                w = torch.cat([self.gblock(name, 0, 0), self.gblock(name, 1, 0)], dim=2)
                out[name] = (w[:, :, :Cc].cpu().numpy(), self.gbias(name, 0).cpu().numpy())
            elif k == 1:
                w = torch.cat([self.gblock(name, 0, 0), self.gblock(name, 0, 1)], dim=1)
                b = torch.cat([self.gbias(name, 0), self.gbias(name, 1)])
                out[name] = (w[:, :Cc].reshape(1, 1, 3, Cc).cpu().numpy(), b[:Cc].cpu().numpy())
            else:
                out[name] = (self.gblock(name, 0, 0).cpu().numpy(), self.gbias(name, 0).cpu().numpy())
        return out

    def repack_t(self):
        net = self.net
        net.run_pack_table(self._pack_table_t)
        for i in range(2):      # tail dgrad as a 1x1 conv over the im2col'ed gradient: B[j = (ky,kx,co)][ci] = W_i[ky][kx][ci][co]
            w = net.blocks[(net.tail, i, 0)]
            self.tail_colw[i].view(NUMK, NUMK)[:27].copy_(w.permute(0, 1, 3, 2).reshape(27, NUMK))
            L.check(self.lib.sr_pack_conv_weights(L.ptr(self.tail_colw[i]), 1, NUMK, 0, L.ptr(self.tail_colw_packed[i]),
                                                  L.stream_ptr()))

    def graph(self, NB, H, W):
        key = (NB, H, W)
        g = self._graphs.pop(key, None)
        if g is None:
            while len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            g = _PlaneTrainGraph(self, NB, H, W)
        self._graphs[key] = g
        return g

    def _load(self, g, x, y):
        for dst, src in ((g.x_in, x), (g.y_true, y)):
            if isinstance(src, torch.Tensor):
                dst.copy_(src, non_blocking=True)
            else:
                dst.copy_(torch.from_numpy(np.ascontiguousarray(src, dtype=np.float32)), non_blocking=True)

    def forward_backward_device(self, g):
        """Forward + backward on the tensors in g.x_in / g.y_true; gradients land in self.grads.  A fixed launch
        sequence on fixed buffers: eager once, then captured and replayed as one CUDA graph (the 192-channel graph at
        training sizes is ~1400 short launches)."""
        if g.cuda_graph is not None:
            g.cuda_graph.replay()
            return

        def body():
            st = L.stream_ptr()
            self.grads.zero_()
            g.loss_sum.zero_()
            for f in g.fwd:
                f(st)
            for f in g.bwd:
                f(st)

        body()
        if self.net.use_graphs and g.ran_eager and not torch.cuda.is_current_stream_capturing():
            try:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    body()
                g.cuda_graph = gr
            except Exception as e:  # noqa: BLE001  (capture unsupported here: stay eager, but say so)
                import warnings
                warnings.warn("sr100: CUDA graph capture of the training step failed (%s: %s); running eagerly"
                              % (type(e).__name__, e), RuntimeWarning)
                self.net.use_graphs = False
        g.ran_eager = True

    def apply_gradients(self):
        from .dist import all_reduce_sum_
        world = all_reduce_sum_(self.grads)
        self.t += 1
        L.check(self.lib.sr_adam_step(L.ptr(self.net.param_arena), L.ptr(self.grads), L.ptr(self.m), L.ptr(self.v),
                                      self.net.n_arena, self.lr, self.beta_1, self.beta_2, self.epsilon, self.t,
                                      1.0 / world, L.stream_ptr()))
        self.net.repack()
        self.repack_t()

    def step_device(self, g):
        self.forward_backward_device(g)
        self.apply_gradients()

    def train_on_batch(self, x, y):
        xs, s = tuple(x.shape), self.net.scale
        if len(xs) != 4 or xs[-1] != 3 or tuple(y.shape) != (xs[0], s * xs[1], s * xs[2], 3):
            raise ValueError("Error when checking target: expected x (N,h,w,3) and y (N,%dh,%dw,3), got %s and %s"
                             % (s, s, xs, tuple(y.shape)))
        g = self.graph(*xs[:3])
        self._load(g, x, y)
        self.step_device(g)
        return self.last_loss(g)

    def last_loss(self, g):
        from .dist import all_reduce_sum_
        t = torch.stack([g.loss_sum[0], torch.tensor(float(g.n_local), dtype=torch.float64, device=g.loss_sum.device)])
        all_reduce_sum_(t)
        sse, n = t.tolist()
        return sse / n

    def evaluate(self, x, y):
        """(mse, categorical accuracy over the 3 colour channels) -- the compile(metrics=['accuracy']) pair."""
        g = self.graph(*tuple(x.shape)[:3])
        self._load(g, x, y)
        st = L.stream_ptr()
        for f in g.fwd:
            f(st)
        d = g.out - g.y_true
        acc = (g.out.argmax(dim=-1) == g.y_true.argmax(dim=-1)).float().mean()
        return float((d * d).mean().item()), float(acc.item())
