"""Forward engine of the DifvdsrDouble x4 stack on sm_100a.

Executes the graph of models.DifvdsrDouble.create_model (reference models.py:1159-1270) as a fixed
sequence of libsr100 launches on device-resident NHWC tensors:

    head 1x1 (CUDA cores) -> 16 x [conv3+relu, conv5+relu, (conv5 (+) conv3) fused 0.1/0.9 residual]
    -> 6 x [conv3+relu, conv3 fused 0.1/1.0 residual] -> bilinear x4 (TF1 legacy) -> 2 x 5/3 block at HR
    -> tail conv3 -> relu (fp32 [N,4H,4W,3]).

All 128-channel convs run on tcgen05 tensor cores (bf16 operands, fp32 TMEM accumulators).  The
residual stream is kept in fp32 in HBM at LR (22 blocks of accumulation) and in bf16 at HR (2 blocks);
`stream` chooses otherwise.  PyTorch is only the allocator / stream provider here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L

NUMK = 128


def layer_specs():
    """[(name, ksize, cin, cout)] in Keras creation order (models.py:1177-1199): 'level1', then
    conv2d_1..conv2d_85 (k3,k5,k5,k3 per 5/3 block; k3,k3 per light block; tail last)."""
    specs = [("level1", 1, 3, NUMK)]
    n = 0
    for _ in range(16):
        for k in (3, 5, 5, 3):
            n += 1
            specs.append(("conv2d_%d" % n, k, NUMK, NUMK))
    for _ in range(6):
        for k in (3, 3):
            n += 1
            specs.append(("conv2d_%d" % n, k, NUMK, NUMK))
    for _ in range(2):
        for k in (3, 5, 5, 3):
            n += 1
            specs.append(("conv2d_%d" % n, k, NUMK, NUMK))
    n += 1
    specs.append(("conv2d_%d" % n, 3, NUMK, 3))
    return specs


def glorot_uniform_weights(seed=None):
    """Keras default initialisation: glorot_uniform kernels, zero biases."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, k, cin, cout in layer_specs():
        limit = np.sqrt(6.0 / (k * k * cin + k * k * cout))
        out[name] = (rng.uniform(-limit, limit, size=(k, k, cin, cout)).astype(np.float32),
                     np.zeros((cout,), dtype=np.float32))
    return out


class _Plan:
    """Owning wrapper of one sr_conv_plan."""

    def __init__(self, lib, desc):
        self.lib = lib
        self.handle = C.c_void_p()
        L.check(lib.sr_conv_plan_create(C.byref(desc), C.byref(self.handle)))
        info = L.ConvPlanInfo()
        L.check(lib.sr_conv_plan_info(self.handle, C.byref(info)))
        self.flops = info.flops
        self.info = info

    def run(self, stream):
        L.check(self.lib.sr_conv_plan_run(self.handle, stream))

    def __del__(self):
        try:
            if self.handle:
                self.lib.sr_conv_plan_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


class _Graph:
    """Buffers + conv plans for one (NB, H, W) input shape."""

    def __init__(self, eng, NB, H, W):
        dev = eng.device
        bf, f32 = torch.bfloat16, torch.float32
        self.eng = eng
        self.cuda_graph, self.ran_eager = None, False
        self.NB, self.H, self.W = NB, H, W
        self.x_in = torch.empty(NB, H, W, 3, device=dev, dtype=f32)
        self.s_lr = torch.empty(NB, H, W, NUMK, device=dev, dtype=bf)
        self.t1_lr = torch.empty_like(self.s_lr)
        self.t2_lr = torch.empty_like(self.s_lr)
        self.s_lr32 = torch.empty(NB, H, W, NUMK, device=dev, dtype=f32) if eng.stream_lr_fp32 else None
        HH, WW = 4 * H, 4 * W
        self.s_hr = torch.empty(NB, HH, WW, NUMK, device=dev, dtype=bf)
        self.t1_hr = torch.empty_like(self.s_hr)
        self.t2_hr = torch.empty_like(self.s_hr)
        self.s_hr32 = torch.empty(NB, HH, WW, NUMK, device=dev, dtype=f32) if eng.stream_hr_fp32 else None
        self.out = torch.empty(NB, HH, WW, 3, device=dev, dtype=f32)
        self.steps = []  # list of callables(stream)
        self.conv_flops = 0.0
        lib = eng.lib
        names = [s[0] for s in layer_specs()]

        def conv(srcs, out_bf16=None, out_f32=None, relu=0, alpha=1.0, beta=0.0, res32=None, res16=None,
                 shape=(NB, H, W), cout=NUMK):
            d = L.ConvDesc()
            d.nsrc = len(srcs)
            bias = None
            for s, (name, x) in enumerate(srcs):
                d.in_[s] = x.data_ptr()
                d.wpacked[s] = eng.packed[name].data_ptr()
                d.ksize[s] = eng.ksize[name]
            bias = eng.bias_for(tuple(n for n, _ in srcs))
            d.NB, d.H, d.W = shape
            d.cin, d.cout = NUMK, cout
            d.bias = bias.data_ptr()
            d.alpha, d.beta, d.relu = alpha, beta, relu
            d.res_f32 = res32.data_ptr() if res32 is not None else None
            d.res_bf16 = res16.data_ptr() if (res16 is not None and res32 is None) else None
            d.out_bf16 = out_bf16.data_ptr() if out_bf16 is not None else None
            d.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
            d.a_mode, d.nacc, d.pair = eng.a_mode, eng.nacc, eng.pair
            p = _Plan(lib, d)
            self.conv_flops += p.flops
            self.steps.append(p.run)
            return p

        def block53(i, s, s32, t1, t2, shape):
            conv([(names[i], s)], out_bf16=t1, relu=1, shape=shape)
            conv([(names[i + 2], s)], out_bf16=t2, relu=1, shape=shape)
            conv([(names[i + 1], t1), (names[i + 3], t2)], out_bf16=s, out_f32=s32, alpha=0.1, beta=0.9,
                 res32=s32, res16=s, shape=shape)

        def block_light(i, s, s32, t1, shape):
            conv([(names[i], s)], out_bf16=t1, relu=1, shape=shape)
            conv([(names[i + 1], t1)], out_bf16=s, out_f32=s32, alpha=0.1, beta=1.0, res32=s32, res16=s,
                 shape=shape)

        npix = NB * H * W
        w0, b0 = eng.head_w, eng.head_b
        self.steps.append(lambda st: L.check(lib.sr_head1x1_fwd(
            L.ptr(self.x_in), L.ptr(w0), L.ptr(b0), npix, L.ptr(self.s_lr), L.ptr(self.s_lr32), st)))
        i = 1
        lr = (NB, H, W)
        for _ in range(16):
            block53(i, self.s_lr, self.s_lr32, self.t1_lr, self.t2_lr, lr)
            i += 4
        for _ in range(6):
            block_light(i, self.s_lr, self.s_lr32, self.t1_lr, lr)
            i += 2
        src = self.s_lr32 if self.s_lr32 is not None else self.s_lr
        src_is_bf16 = 0 if self.s_lr32 is not None else 1
        self.steps.append(lambda st: L.check(lib.sr_bilinear4_fwd(
            L.ptr(src), src_is_bf16, NB, H, W, NUMK, L.ptr(self.s_hr), L.ptr(self.s_hr32), st)))
        hr = (NB, HH, WW)
        for _ in range(2):
            block53(i, self.s_hr, self.s_hr32, self.t1_hr, self.t2_hr, hr)
            i += 4
        conv([(names[i], self.s_hr)], out_f32=self.out, relu=1, shape=hr, cout=3)

    def run(self):
        """One forward over the resident x_in.  The launch sequence is fixed (plans are bound to these buffers), so
        after one eager run it is captured into a CUDA graph and replayed: one submission instead of ~70 ctypes
        launches (matters for small inputs such as a single 128x128 patch, where launches cost as much as math)."""
        eng = self.eng
        if eng.use_graphs and self.cuda_graph is not None:
            self.cuda_graph.replay()
            return self.out
        st = L.stream_ptr()
        for step in self.steps:
            step(st)
        if eng.use_graphs and self.cuda_graph is None and self.ran_eager and not torch.cuda.is_current_stream_capturing():
            try:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    cst = L.stream_ptr()
                    for step in self.steps:
                        step(cst)
                self.cuda_graph = gr
            except Exception:  # noqa: BLE001  (capture unsupported in this context: stay eager)
                eng.use_graphs = False
        self.ran_eager = True
        return self.out


class Engine:
    """Device-resident DifvdsrDouble weights + cached per-shape graphs."""

    def __init__(self, weights=None, device=None, stream="lr32", a_mode=0, nacc=2, pair=1,
                 max_pixels=192 * 96 * 96, use_graphs=True):
        self.lib = L.require_device()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        assert stream in ("bf16", "lr32", "fp32")
        self.stream_lr_fp32 = stream in ("lr32", "fp32")
        self.stream_hr_fp32 = stream == "fp32"
        self.a_mode, self.nacc, self.pair = a_mode, nacc, pair
        self.max_pixels = max_pixels  # LR pixels per sub-batch (HR activations are 16x this)
        self.use_graphs = use_graphs and os.environ.get("SR100_NO_GRAPHS", "0") != "1"
        # (weights, biases and activations are read through fixed device pointers: set_weights / repack rewrite them
        #  in place, so captured graphs stay valid)
        self.specs = layer_specs()
        self.ksize = {n: k for n, k, _, _ in self.specs}
        self.master = {}   # name -> (kernel HWIO fp32 device, bias fp32 device): views into param_arena
        # one flat fp32 arena in layer order (kernel, bias, kernel, bias, ...): the optimizer step and the
        # gradient all-reduce of the training path are single launches over it
        self.param_slices = {}
        off = 0
        for name, k, cin, cout in self.specs:
            nw, nb = k * k * cin * cout, cout
            self.param_slices[name] = (off, nw, off + nw, nb)
            off += nw + nb
        self.n_params = off
        self.param_arena = torch.zeros(off, dtype=torch.float32, device=self.device)
        for name, k, cin, cout in self.specs:
            ow, nw, ob, nb = self.param_slices[name]
            self.master[name] = (self.param_arena[ow:ow + nw].view(k, k, cin, cout), self.param_arena[ob:ob + nb])
        self.packed = {}
        self._bias_cache = {}
        self._graphs = {}
        self.set_weights_dict(weights if weights is not None else glorot_uniform_weights())

    # ---------------------------------------------------------------- weights
    def set_weights_dict(self, weights):
        st = L.stream_ptr()
        for name, k, cin, cout in self.specs:
            w, b = weights[name]
            w = np.ascontiguousarray(w, dtype=np.float32)
            b = np.ascontiguousarray(b, dtype=np.float32)
            if w.shape != (k, k, cin, cout) or b.shape != (cout,):
                raise ValueError("layer %s: expected kernel %s / bias %s, got %s / %s"
                                 % (name, (k, k, cin, cout), (cout,), w.shape, b.shape))
            self.master[name][0].copy_(torch.from_numpy(w))
            self.master[name][1].copy_(torch.from_numpy(b))
        self.head_w = self.master["level1"][0].reshape(3, NUMK)
        self.head_b = self.master["level1"][1]
        self.repack()

    def repack(self):
        """(Re)build the tensor-core weight layout from the fp32 masters (after load / optimizer step)."""
        st = L.stream_ptr()
        for name, k, cin, cout in self.specs:
            if cin != NUMK:
                continue
            if name not in self.packed:
                self.packed[name] = torch.empty(self.lib.sr_packed_weight_bytes(k, cout), dtype=torch.uint8,
                                                device=self.device)
            L.check(self.lib.sr_pack_conv_weights(L.ptr(self.master[name][0]), k, cout, 0,
                                                  L.ptr(self.packed[name]), st))
        for names, t in self._bias_cache.items():
            t.copy_(sum(self.master[n][1] for n in names))

    def bias_for(self, names):
        if names not in self._bias_cache:
            self._bias_cache[names] = sum(self.master[n][1] for n in names).clone()
        return self._bias_cache[names]

    def get_weights_dict(self):
        return {n: (self.master[n][0].cpu().numpy(), self.master[n][1].cpu().numpy()) for n, _, _, _ in self.specs}

    # ---------------------------------------------------------------- forward
    def graph(self, NB, H, W):
        key = (NB, H, W)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 4:  # bound device memory: keep the most recent shapes only
                self._graphs.pop(next(iter(self._graphs)))
            g = _Graph(self, NB, H, W)
            self._graphs[key] = g
        return g

    def sub_batch(self, H, W):
        return max(1, self.max_pixels // (H * W))

    def forward_device(self, x, out=None):
        """x: device float32 [N,H,W,3] in [0,1] -> device float32 [N,4H,4W,3] (model.predict)."""
        N, H, W, _ = x.shape
        if out is None:
            out = torch.empty(N, 4 * H, 4 * W, 3, device=self.device, dtype=torch.float32)
        nb = min(N, self.sub_batch(H, W))
        for i in range(0, N, nb):
            n = min(nb, N - i)
            g = self.graph(n, H, W)
            g.x_in.copy_(x[i:i + n])
            out[i:i + n].copy_(g.run())
        return out

    def upscale_images_device(self, imgs_u8, patch=96, step=64, scale=4):
        """The device part of upscaleStepPatch (models.py:225-391) for a list of uint8 [h,w,3] device images:
        zero-padded canvas -> 96/64 patch gather (/255) -> conv stack over ALL tiles of all images -> x255,
        stitch with the 8-px crop, clip -> uint8.  Returns the uncropped uint8 canvases (device)."""
        from . import ops
        metas, parts = [], []
        for img in imgs_u8:
            h, w, _ = img.shape
            ch, cw = ops.canvas_size(h, w, patch, step)
            p, counts = ops.patch_gather_u8(img, (ch, cw), (patch, patch), step, divisor=255.0)
            metas.append((ch, cw, counts, p.shape[0]))
            parts.append(p)
        allp = parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
        out = self.forward_device(allp)
        res, off = [], 0
        for ch, cw, counts, n in metas:
            _, u8 = ops.patch_stitch(out[off:off + n], counts, (patch, patch), step, scale, (ch, cw), mul=255.0,
                                     want_f32=False, want_u8=True)
            res.append(u8)
            off += n
        return res

    def conv_flops(self, N, H, W):
        """Algorithmic FLOPs (2*MAC) of one forward over N patches of HxW (tensor-core convs + head)."""
        lr = N * H * W
        per_lr = 16 * (2 * 9 + 2 * 25) * NUMK * NUMK * 2 + 6 * 2 * 9 * NUMK * NUMK * 2 + 2 * 3 * NUMK
        per_hr = 2 * (2 * 9 + 2 * 25) * NUMK * NUMK * 2 + 9 * NUMK * 3 * 2
        return float(lr) * per_lr + float(lr) * 16 * per_hr
