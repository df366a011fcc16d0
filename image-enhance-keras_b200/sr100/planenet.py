"""Forward engines of the reference's two older graphs on sm_100a: Difvdsr4 (256 channels, bilinear x2 twice,
models.py:992-1142) and Difvdsr (192 channels, same resolution, "difference" residual blocks,
models.py:1274-1357).

The tensor-core conv kernel is a 128 -> 128 channel kernel that can accumulate TWO sources into one output
(`sr_conv_desc.nsrc = 2`, written for the fused 5/3 block tail of DifvdsrDouble).  A C-channel activation
(C = 256 or 192) is therefore kept as TWO PLANES of [NB,H,W,128] (bf16 operand copy + fp32 residual stream; the
second plane of a 192-channel tensor carries zeros in channels 64..127), and a C -> C convolution is two launches
(one per output plane), each accumulating both input planes with the matching 128 x 128 slices of the Keras HWIO
kernel.  Everything the blocks do besides the contraction is epilogue algebra of those launches:

    light0 / light block (Difvdsr4)   t = act(conv(x))                       act = LeakyReLU(0.001) / ReLU
                                      x = 0.1 * conv(t) + x                   alpha = 0.1, beta = 1, res = x (fp32)
    difference block (Difvdsr)        t = relu(conv1(x))
                                      d = conv2(t) - x                        alpha = 1, beta = -1, res = x
                                      r = 1.1 * x + 0.2 * d                   (sr_axpby_f32)
                                      u = LeakyReLU(0.2)(conv3(d))
                                      x = 0.1 * conv4(u) + r                  = 0.1*(d + e + a) + x with a = d + x
The reference's `y = 0.1*(d + e + a) + x` (models.py:1352-1355) is evaluated as 0.1*e + 0.2*d + 1.1*x: the same value
up to fp32 rounding order.  Heads: Difvdsr4 1x1 (sr_head1x1_fwd per plane), Difvdsr 3x3 from 3 channels
(sr_conv2d_direct); tails: the 3-output tensor-core conv with both planes as sources.  No CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from .engine import _Plan, _Stage

NUMK = 128


def difvdsr4_specs(numk=256):
    """[(name, k, cin, cout)] in Keras creation order (models.py:1024-1047)."""
    specs = [("level1", 1, 3, numk)]
    for i in range(1, 2 * (6 + 20 + 6) + 1):
        specs.append(("conv2d_%d" % i, 3, numk, numk))
    specs.append(("conv2d_%d" % len(specs), 3, numk, 3))
    return specs


def difvdsr_specs(numk=192):
    """models.py:1304-1308."""
    specs = [("level1", 3, 3, numk)]
    for i in range(1, 4 * 32 + 1):
        specs.append(("conv2d_%d" % i, 3, numk, numk))
    specs.append(("conv2d_%d" % len(specs), 3, numk, 3))
    return specs


ARCHS = {"difvdsr4": (difvdsr4_specs, 256, 4), "difvdsr": (difvdsr_specs, 192, 1)}


def glorot_uniform(specs, seed=None):
    rng = np.random.default_rng(seed)
    out = {}
    for name, k, cin, cout in specs:
        limit = np.sqrt(6.0 / (k * k * cin + k * k * cout))
        out[name] = (rng.uniform(-limit, limit, size=(k, k, cin, cout)).astype(np.float32),
                     np.zeros((cout,), dtype=np.float32))
    return out


class _Net(_Stage):
    """Buffers + launch list of one (NB, H, W) input shape."""

    def __init__(self, eng, NB, H, W):
        self._init_stage(eng)
        self.NB, self.H, self.W = NB, H, W
        dev = eng.device
        self.x_in = torch.empty(NB, H, W, 3, device=dev, dtype=torch.float32)
        s = eng.scale
        self.out = torch.empty(NB, s * H, s * W, 3, device=dev, dtype=torch.float32)
        self.keep = []
        (self._build_difvdsr4 if eng.arch == "difvdsr4" else self._build_difvdsr)()

    # ------------------------------------------------------------------ helpers
    def planes(self, shape, dtype, zero=False):
        mk = torch.zeros if zero else torch.empty
        t = [mk(*shape, NUMK, device=self.eng.device, dtype=dtype) for _ in range(2)]
        self.keep.append(t)
        return t

    def conv(self, name, src16, shape, out16=None, out32=None, relu=0, slope=0.0, alpha=1.0, beta=0.0, res32=None):
        eng = self.eng
        for j in range(2):
            d = L.ConvDesc()
            d.nsrc = 2
            for i in range(2):
                d.in_[i] = src16[i].data_ptr()
                d.wpacked[i] = eng.packed[(name, i, j)].data_ptr()
                d.ksize[i] = 3
            d.NB, d.H, d.W = shape
            d.cin, d.cout = NUMK, NUMK
            d.cin_valid[1] = eng.C - NUMK if eng.skip_zero_k else 0    # 192 channels: plane 1 carries 64 real ones
            d.bias = eng.bias[(name, j)].data_ptr()
            d.alpha, d.beta, d.relu, d.leaky_slope = alpha, beta, relu, slope
            d.res_f32 = res32[j].data_ptr() if res32 is not None else None
            d.out_bf16 = out16[j].data_ptr() if out16 is not None else None
            d.out_f32 = out32[j].data_ptr() if out32 is not None else None
            d.a_mode, d.nacc, d.pair = 0, 2, 1
            p = _Plan(eng.lib, d)
            self.conv_flops += p.flops
            self.steps.append(p.run)

    def tail(self, name, src16, shape):
        eng = self.eng
        d = L.ConvDesc()
        d.nsrc = 2
        for i in range(2):
            d.in_[i] = src16[i].data_ptr()
            d.wpacked[i] = eng.packed[(name, i, 0)].data_ptr()
            d.ksize[i] = 3
        d.NB, d.H, d.W = shape
        d.cin, d.cout = NUMK, 3
        d.cin_valid[1] = eng.C - NUMK if eng.skip_zero_k else 0
        d.bias = eng.bias[(name, 0)].data_ptr()
        d.alpha, d.beta, d.relu = 1.0, 0.0, 1
        d.out_f32 = self.out.data_ptr()
        d.a_mode, d.nacc, d.pair = 0, 2, 0
        p = _Plan(eng.lib, d)
        self.conv_flops += p.flops
        self.steps.append(p.run)

    def axpby(self, x32, y32, a, b, out32, out16=None):
        lib = self.eng.lib
        for j in range(2):
            n = x32[j].numel()
            self.steps.append(lambda st, j=j, n=n: L.check(lib.sr_axpby_f32(
                L.ptr(x32[j]), L.ptr(y32[j]), a, b, n, L.ptr(out32[j]) if out32 is not None else None,
                L.ptr(out16[j]) if out16 is not None else None, st)))

    def light_block(self, i, s16, s32, t16, shape, leaky=None):
        names = self.eng.names
        if leaky is None:
            self.conv(names[i], s16, shape, out16=t16, relu=1)
        else:
            self.conv(names[i], s16, shape, out16=t16, relu=2, slope=leaky)
        self.conv(names[i + 1], t16, shape, out16=s16, out32=s32, alpha=0.1, beta=1.0, res32=s32)

    def upsample2(self, s32, shape):
        NB, H, W = shape
        up = (NB, 2 * H, 2 * W)
        o16, o32 = self.planes(up, torch.bfloat16), self.planes(up, torch.float32)
        lib = self.eng.lib
        for j in range(2):
            self.steps.append(lambda st, j=j: L.check(lib.sr_bilinear2_fwd(
                L.ptr(s32[j]), 0, NB, H, W, NUMK, L.ptr(o16[j]), L.ptr(o32[j]), st)))
        return o16, o32, up

    # ------------------------------------------------------------------ graphs
    def _build_difvdsr4(self):
        eng, lib = self.eng, self.eng.lib
        bf, f32 = torch.bfloat16, torch.float32
        shape = (self.NB, self.H, self.W)
        npix = self.NB * self.H * self.W
        s16, s32, t16 = self.planes(shape, bf), self.planes(shape, f32), self.planes(shape, bf)
        for j in range(2):                                      # level1: 1x1, 3 -> 256, ReLU (models.py:1024)
            self.steps.append(lambda st, j=j, o16=s16[j], o32=s32[j]: L.check(lib.sr_head1x1_fwd(
                L.ptr(self.x_in), L.ptr(eng.head_w[j]), L.ptr(eng.bias[("level1", j)]), npix, L.ptr(o16),
                L.ptr(o32), st)))
        i = 1
        for _ in range(6):                                      # :1030-1032
            self.light_block(i, s16, s32, t16, shape, leaky=0.001)
            i += 2
        s16, s32, shape = self.upsample2(s32, shape)            # :1034
        t16 = self.planes(shape, bf)
        xinp = self.planes(shape, f32)                          # xInp = x (:1035)
        for j in range(2):
            self.steps.append(lambda st, dst=xinp[j], src=s32[j]: dst.copy_(src))
        for _ in range(20):                                     # :1036-1038
            self.light_block(i, s16, s32, t16, shape)
            i += 2
        self.axpby(s32, xinp, 1.0, 1.0, s32, s16)               # Add([x, xInp]) (:1039)
        s16, s32, shape = self.upsample2(s32, shape)            # :1041
        t16 = self.planes(shape, bf)
        for _ in range(6):                                      # :1042-1044
            self.light_block(i, s16, s32, t16, shape)
            i += 2
        self.tail(eng.names[i], s16, shape)                     # :1047

    def _build_difvdsr(self):
        eng, lib = self.eng, self.eng.lib
        bf, f32 = torch.bfloat16, torch.float32
        NB, H, W = self.NB, self.H, self.W
        shape = (NB, H, W)
        C = eng.C
        s16, s32 = self.planes(shape, bf, zero=True), self.planes(shape, f32, zero=True)
        t16, d16, d32, r32 = self.planes(shape, bf), self.planes(shape, bf), self.planes(shape, f32), self.planes(shape, f32)
        head = torch.empty(NB, H, W, C, device=eng.device, dtype=f32)
        self.keep.append(head)
        w0, b0 = eng.blocks[("level1", 0, 0)], eng.bias[("level1", 0)]
        self.steps.append(lambda st: L.check(lib.sr_conv2d_direct(          # level1: 3x3, 3 -> 192, ReLU (:1304)
            L.ptr(self.x_in), 0, L.ptr(w0), 0, L.ptr(b0), NB, H, W, 3, C, 3, 1, 1, 0, 0, L.ptr(head), st)))

        def split(st):
            s32[0].copy_(head[..., :NUMK])
            s32[1][..., :C - NUMK].copy_(head[..., NUMK:])
            s16[0].copy_(s32[0])
            s16[1].copy_(s32[1])
        self.steps.append(split)
        names = eng.names
        i = 1
        for _ in range(32):                                                  # :1305-1306, body :1336-1357
            self.conv(names[i], s16, shape, out16=t16, relu=1)
            self.conv(names[i + 1], t16, shape, out16=d16, out32=d32, alpha=1.0, beta=-1.0, res32=s32)
            self.axpby(s32, d32, 1.1, 0.2, r32)
            self.conv(names[i + 2], d16, shape, out16=t16, relu=2, slope=0.2)
            self.conv(names[i + 3], t16, shape, out16=s16, out32=s32, alpha=0.1, beta=1.0, res32=r32)
            i += 4
        self.tail(names[i], s16, shape)                                      # :1308


class PlaneNet:
    """Device-resident weights of Difvdsr4 / Difvdsr + cached per-shape launch lists.

    Parameters live in ONE flat fp32 arena, stored the way the kernels consume them: a C -> C kernel as its four
    dense [k,k,128,128] plane blocks (input plane i, output plane j; rows / columns beyond C are zero and stay zero),
    a bias as two [128] blocks, the Difvdsr4 head as two [3,128] matrices, the tail as two [k,k,128,3] blocks; only
    the 3-channel Difvdsr head stays a dense Keras kernel (it runs on the direct conv).  Packing for the tensor
    cores is then one batched launch over dense blocks, and training (sr100.planetrain) writes filter gradients
    straight into a gradient arena of the same layout and updates everything with one Adam launch.
    get_weights_dict / set_weights_dict convert to and from the Keras HWIO tensors."""

    def __init__(self, arch, weights=None, device=None, use_graphs=True, max_cached=4):
        import os
        if arch not in ARCHS:
            raise ValueError("unknown architecture %r" % (arch,))
        self.lib = L.require_device()
        self.arch = arch
        spec_fn, self.C, self.scale = ARCHS[arch]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.use_graphs = use_graphs and os.environ.get("SR100_NO_GRAPHS", "0") != "1"
        # the zero half of a 192-channel tensor's second plane is not multiplied (sr_conv_desc.cin_valid); the switch
        # exists for the bit-for-bit test against the padded launches
        self.skip_zero_k = self.C < 2 * NUMK and os.environ.get("SR100_SKIP_ZERO_K", "1") != "0"
        self.specs = spec_fn()
        self.names = [s[0] for s in self.specs]
        self.ksize = {s[0]: s[1] for s in self.specs}
        # ---- arena layout: name -> [(kind, key, offset, shape)]
        self.layout, off = {}, 0

        def add(name, kind, key, shape):
            nonlocal off
            n = int(np.prod(shape))
            self.layout.setdefault(name, []).append((kind, key, off, tuple(shape)))
            off += (n + 3) // 4 * 4                           # every piece 16-byte aligned
        for name, k, cin, cout in self.specs:
            if cin == self.C and cout == self.C:
                for i in range(2):
                    for j in range(2):
                        add(name, "w", (i, j), (k, k, NUMK, NUMK))
                for j in range(2):
                    add(name, "b", j, (NUMK,))
            elif cin == self.C:                              # tail: C -> 3
                for i in range(2):
                    add(name, "w", (i, 0), (k, k, NUMK, cout))
                add(name, "b", 0, (cout,))
            elif k == 1:                                     # Difvdsr4 head: [1,1,3,256] -> two [3,128] matrices
                for j in range(2):
                    add(name, "w", (0, j), (3, NUMK))
                for j in range(2):
                    add(name, "b", j, (NUMK,))
            else:                                            # Difvdsr head: dense 3x3x3xC kernel for the direct conv
                add(name, "w", (0, 0), (k, k, cin, cout))
                add(name, "b", 0, (cout,))
        self.n_arena = off
        self.param_arena = torch.zeros(off, dtype=torch.float32, device=self.device)
        self.blocks, self.bias, self.packed = {}, {}, {}
        for name, pieces in self.layout.items():
            for kind, key, o, shape in pieces:
                view = self.param_arena[o:o + int(np.prod(shape))].view(*shape)
                if kind == "w":
                    self.blocks[(name,) + key] = view
                else:
                    self.bias[(name, key)] = view
        self.head_w = [self.blocks[("level1", 0, j)] for j in range(2)] if self.ksize["level1"] == 1 else None
        self.tail = self.names[-1]
        for name, k, cin, cout in self.specs:
            if cin == self.C:
                pc = NUMK if cout == self.C else cout
                for i in range(2):
                    for j in range(2 if cout == self.C else 1):
                        self.packed[(name, i, j)] = torch.empty(self.lib.sr_packed_weight_bytes(k, pc),
                                                                dtype=torch.uint8, device=self.device)
        self._pack_table = self.make_pack_table(self.packed, flip=False)
        self._nets = {}
        self.max_cached = max_cached
        self.set_weights_dict(weights if weights is not None else glorot_uniform(self.specs))

    @property
    def n_params(self):
        return sum(k * k * ci * co + co for _, k, ci, co in self.specs)

    # ------------------------------------------------------------------ packing
    def make_pack_table(self, packed, flip):
        """(items, starts, n, total) of one sr_pack_conv_weights_batched launch over every plane block that has a
        buffer in `packed`: packed[(name, a, b)] = block(name, a, b) as is (flip False), or -- for the input-gradient
        convs -- packed[(name, j, i)] = block(name, i, j) rotated by 180 degrees with cin <-> cout swapped."""
        items, starts, total = [], [0], 0
        for (name, a, b), buf in packed.items():
            blk = self.blocks[(name, b, a)] if flip else self.blocks[(name, a, b)]
            it = L.PackItem()
            it.hwio, it.dst = blk.data_ptr(), buf.data_ptr()
            it.ksize, it.cout, it.transpose_flip = self.ksize[name], blk.shape[3], 1 if flip else 0
            items.append(it)
            total += buf.numel() // 2
            starts.append(total)
        arr = (L.PackItem * len(items))(*items)
        raw = np.frombuffer(memoryview(arr), dtype=np.uint8).copy()
        return (torch.from_numpy(raw).to(self.device), torch.tensor(starts, dtype=torch.int64).to(self.device),
                len(items), total)

    def run_pack_table(self, table):
        items, starts, n, total = table
        L.check(self.lib.sr_pack_conv_weights_batched(L.ptr(items), L.ptr(starts), n, total, L.stream_ptr()))

    def repack(self):
        """Tensor-core layouts of every plane block from the current arena: one launch."""
        self.run_pack_table(self._pack_table)

    # ------------------------------------------------------------------ Keras-shaped weights
    @property
    def master(self):
        """{name: (kernel HWIO, bias)} as fresh device tensors assembled from the arena (read-only view of the
        weights for get_weights / the Keras facade)."""
        out = {}
        C = self.C
        for name, k, cin, cout in self.specs:
            if cin == C and cout == C:
                w = torch.zeros(k, k, 2 * NUMK, 2 * NUMK, device=self.device)
                for i in range(2):
                    for j in range(2):
                        w[:, :, i * NUMK:(i + 1) * NUMK, j * NUMK:(j + 1) * NUMK] = self.blocks[(name, i, j)]
                b = torch.cat([self.bias[(name, 0)], self.bias[(name, 1)]])
                out[name] = (w[:, :, :C, :C].contiguous(), b[:C].contiguous())
            elif cin == C:
                w = torch.cat([self.blocks[(name, 0, 0)], self.blocks[(name, 1, 0)]], dim=2)
                out[name] = (w[:, :, :C].contiguous(), self.bias[(name, 0)].clone())
            elif k == 1:
                w = torch.cat([self.blocks[(name, 0, 0)], self.blocks[(name, 0, 1)]], dim=1)
                b = torch.cat([self.bias[(name, 0)], self.bias[(name, 1)]])
                out[name] = (w[:, :C].reshape(1, 1, 3, C).contiguous(), b[:C].contiguous())
            else:
                out[name] = (self.blocks[(name, 0, 0)].clone(), self.bias[(name, 0)].clone())
        return out

    def get_weights_dict(self):
        return {n: (w.cpu().numpy(), b.cpu().numpy()) for n, (w, b) in self.master.items()}

    def set_weights_dict(self, weights):
        C = self.C
        self.param_arena.zero_()
        for name, k, cin, cout in self.specs:
            w, b = weights[name]
            w = np.ascontiguousarray(w, dtype=np.float32)
            b = np.ascontiguousarray(b, dtype=np.float32)
            if w.shape != (k, k, cin, cout) or b.shape != (cout,):
                raise ValueError("layer %s: expected kernel %s / bias %s, got %s / %s"
                                 % (name, (k, k, cin, cout), (cout,), w.shape, b.shape))
            wd, bd = torch.from_numpy(w).to(self.device), torch.from_numpy(b).to(self.device)
            if cin == C and cout == C:
                for i in range(2):
                    ni = min(NUMK, C - i * NUMK)
                    for j in range(2):
                        nj = min(NUMK, C - j * NUMK)
                        self.blocks[(name, i, j)][:, :, :ni, :nj] = wd[:, :, i * NUMK:i * NUMK + ni, j * NUMK:j * NUMK + nj]
                for j in range(2):
                    nj = min(NUMK, C - j * NUMK)
                    self.bias[(name, j)][:nj] = bd[j * NUMK:j * NUMK + nj]
            elif cin == C:
                for i in range(2):
                    ni = min(NUMK, C - i * NUMK)
                    self.blocks[(name, i, 0)][:, :, :ni] = wd[:, :, i * NUMK:i * NUMK + ni]
                self.bias[(name, 0)].copy_(bd)
            elif k == 1:
                for j in range(2):
                    nj = min(NUMK, C - j * NUMK)
                    self.blocks[(name, 0, j)][:, :nj] = wd.reshape(3, C)[:, j * NUMK:j * NUMK + nj]
                    self.bias[(name, j)][:nj] = bd[j * NUMK:j * NUMK + nj]
            else:
                self.blocks[(name, 0, 0)].copy_(wd)
                self.bias[(name, 0)].copy_(bd)
        self.repack()
        torch.cuda.current_stream().synchronize()

    def net(self, NB, H, W):
        key = (NB, H, W)
        n = self._nets.get(key)
        if n is None:
            while len(self._nets) >= self.max_cached:
                self._nets.pop(next(iter(self._nets)))
            n = _Net(self, NB, H, W)
            self._nets[key] = n
        return n

    def forward_device(self, x):
        """x: float32 [NB,H,W,3] in [0,1] on the device -> float32 [NB, s*H, s*W, 3] (a fresh tensor)."""
        if x.dim() != 4 or x.shape[-1] != 3:
            raise ValueError("expected [NB,H,W,3], got %s" % (tuple(x.shape),))
        NB, H, W, _ = x.shape
        n = self.net(NB, H, W)
        n.x_in.copy_(x)
        n.run()
        return n.out.clone()

    def upscale_images_device(self, imgs, patch=96, step=64, scale=4, full_canvas=False, tiles_per_pass=None):
        """upscaleStepPatch (models.py:184-415) for these graphs, literally: zero-padded canvas, 96/64 tiles,
        /255, predict, x255, 8-px-crop stitch, clip -> uint8; returns the uncropped canvases (full_canvas) or the
        [0, s*h) x [0, s*w) crops (:412).  `scale` is the reference's `scalemulti`: it must equal the graph's own
        factor (4 for Difvdsr4, 1 for Difvdsr) or the reference's stitch would not fit its patches either."""
        from . import ops
        s = self.scale
        if int(scale) != s:
            raise ValueError("scalemulti=%d does not match the x%d output of %s" % (scale, s, self.arch))
        if tiles_per_pass is None:
            tiles_per_pass = 32 if self.arch == "difvdsr4" else 96
        outs = []
        for img in imgs:
            h, w, _ = img.shape
            ch, cw = ops.canvas_size(h, w, patch, step)
            patches, counts = ops.patch_gather_u8(img, (ch, cw), (patch, patch), step, 255.0)
            n = patches.shape[0]
            res = torch.empty(n, patch * s, patch * s, 3, device=self.device, dtype=torch.float32)
            for lo in range(0, n, tiles_per_pass):
                hi = min(n, lo + tiles_per_pass)
                res[lo:hi] = self.forward_device(patches[lo:hi])
            _, u8 = ops.patch_stitch(res, counts, (patch, patch), step, s, (ch, cw), mul=255.0, want_f32=False,
                                     want_u8=True)
            outs.append(u8 if full_canvas else u8[:h * s, :w * s].contiguous())
        return outs
