"""Forward engine of the DifvdsrDouble x4 stack on sm_100a.

Executes the graph of models.DifvdsrDouble.create_model (reference models.py:1159-1270) as a fixed
sequence of libsr100 launches on device-resident NHWC tensors:

    head 1x1 (CUDA cores) -> 16 x [conv3+relu, conv5+relu, (conv5 (+) conv3) fused 0.1/0.9 residual]
    -> 6 x [conv3+relu, conv3 fused 0.1/1.0 residual] -> bilinear x4 (TF1 legacy) -> 2 x 5/3 block at HR
    -> tail conv3 -> relu (fp32 [N,4H,4W,3]).

All 128-channel convs run on tcgen05 tensor cores (bf16 operands, fp32 TMEM accumulators).  The
residual stream is kept in fp32 in HBM at LR (22 blocks of accumulation) and in bf16 at HR (2 blocks);
`stream` chooses otherwise.  precision="tf32" is the accuracy mode (the reference's Conv2D is fp32): every
activation tensor is fp32, operands are rounded to tf32 (round to nearest) by the kernel that produces them, the
MMAs are tcgen05 kind::tf32 and the whole epilogue algebra is fp32 -- the same launch sequence at half the tensor
rate.  PyTorch is only the allocator / stream provider here.
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


class _Stage:
    """A fixed launch sequence on fixed buffers; replayed as a CUDA graph after one eager run (one submission
    instead of tens of ctypes launches: matters for small inputs where launches cost as much as the math)."""

    def _init_stage(self, eng):
        self.eng = eng
        self.steps = []          # callables(stream)
        self.par = set()         # indices i: steps[i] may run concurrently with steps[i - 1] (independent launches)
        self.conv_flops = 0.0    # algorithmic FLOPs (2*MAC) of the tensor-core launches of this stage
        self.cuda_graph, self.ran_eager = None, False

    def _conv(self, srcs, shape, out_bf16=None, out_f32=None, relu=0, alpha=1.0, beta=0.0, res32=None, res16=None,
              cout=NUMK, out_index=None, out_hw=None, comp=None):
        eng = self.eng
        d = L.ConvDesc()
        d.nsrc = len(srcs)
        for s, (name, x) in enumerate(srcs):
            d.in_[s] = x.data_ptr()
            d.wpacked[s] = eng.packed[name].data_ptr()
            d.ksize[s] = eng.ksize[name]
        d.NB, d.H, d.W = shape
        d.cin, d.cout = NUMK, cout
        d.bias = eng.bias_for(tuple(n for n, _ in srcs)).data_ptr()
        d.alpha, d.beta, d.relu = alpha, beta, relu
        d.res_f32 = res32.data_ptr() if res32 is not None else None
        d.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
        if eng.tf32:
            # the "operand copy" slot (bf16 in the default mode) is the tf32-rounded fp32 tensor; residuals are
            # always the fp32 stream
            assert res16 is None or res32 is not None
            d.precision = 1
            d.out_tf32 = out_bf16.data_ptr() if out_bf16 is not None else None
        else:
            d.res_bf16 = res16.data_ptr() if (res16 is not None and res32 is None) else None
            d.out_bf16 = out_bf16.data_ptr() if out_bf16 is not None else None
        if out_index is not None:
            d.out_index, (d.out_h, d.out_w) = out_index.data_ptr(), out_hw
        d.a_mode, d.nacc, d.pair = eng.a_mode, eng.nacc, eng.pair
        if comp is not None:
            d.comp_h, d.comp_w = comp
        p = _Plan(eng.lib, d)
        self.conv_flops += p.flops
        self.steps.append(p.run)
        return p

    def _block53(self, names, i, s, s32, t1, t2, shape, comp=(None, None, None)):
        # comp = compute extents (out, t1, t2): t1 feeds the 5x5 (radius 2), t2 the 3x3 (radius 1)
        pa = self._conv([(names[i], s)], shape, out_bf16=t1, relu=1, comp=comp[1])
        pb = self._conv([(names[i + 2], s)], shape, out_bf16=t2, relu=1, comp=comp[2])
        # the two branch heads read the same tensor and write different ones: when both grids fit on the chip
        # together (small inputs: one CTA per tile, fewer tiles than SMs) they run concurrently on two streams
        if pa.info.grid + pb.info.grid <= self.eng.sm_count:
            self.par.add(len(self.steps) - 1)
        self._conv([(names[i + 1], t1), (names[i + 3], t2)], shape, out_bf16=s, out_f32=s32, alpha=0.1, beta=0.9,
                   res32=s32, res16=s, comp=comp[0])

    def _block_light(self, names, i, s, s32, t1, shape, comp=(None, None)):
        self._conv([(names[i], s)], shape, out_bf16=t1, relu=1, comp=comp[1])
        self._conv([(names[i + 1], t1)], shape, out_bf16=s, out_f32=s32, alpha=0.1, beta=1.0, res32=s32, res16=s,
                   comp=comp[0])

    def run(self):
        eng = self.eng
        if eng.use_graphs and self.cuda_graph is not None:
            self.cuda_graph.replay()
            return
        self._launch_all()
        if eng.use_graphs and self.cuda_graph is None and self.ran_eager and not torch.cuda.is_current_stream_capturing():
            try:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    self._launch_all()
                self.cuda_graph = gr
            except Exception as e:  # noqa: BLE001  (capture unsupported in this context: stay eager, but say so)
                import warnings
                warnings.warn("sr100: CUDA graph capture of a forward stage failed (%s: %s); running eagerly"
                              % (type(e).__name__, e), RuntimeWarning)
                eng.use_graphs = False
        self.ran_eager = True


    def _launch_all(self):
        """Issue the launch sequence on the current stream; steps marked in `par` fork onto the engine's side stream
        and join again (inside a CUDA-graph capture this becomes two parallel branches of the graph)."""
        eng = self.eng
        main = torch.cuda.current_stream()
        st = L.stream_ptr()
        side = eng.side_stream if self.par else None
        i, n = 0, len(self.steps)
        while i < n:
            if side is not None and (i + 1) in self.par:
                fork = torch.cuda.Event()
                fork.record(main)
                side.wait_event(fork)
                self.steps[i + 1](C.c_void_p(side.cuda_stream))
                self.steps[i](st)
                join = torch.cuda.Event()
                join.record(side)
                main.wait_event(join)
                i += 2
            else:
                self.steps[i](st)
                i += 1


class _LRStage(_Stage):
    """Low-resolution stage for NB patches of HxW: head 1x1 + 16 5/3 blocks + 6 light blocks.  Owns the patch
    input buffer, the LR residual stream (the HR stages read it) and the full-size output patch slots."""

    def __init__(self, eng, NB, H, W, need=None):
        """need = (rows, cols) of the final LR stream that the HR stage will read (None: everything).  The layers
        are then restricted, last to first, to the region whose receptive field can still reach it: a light block
        grows the needed region by 2 pixels per axis, a 5/3 block by 3 (lr_extents)."""
        self._init_stage(eng)
        dev, bf, f32 = eng.device, torch.bfloat16, torch.float32
        self.NB, self.H, self.W = NB, H, W
        self.need = (H, W) if need is None else (min(H, need[0]), min(W, need[1]))
        exts = lr_extents(self.need, (H, W))
        op = f32 if eng.tf32 else bf      # dtype of the conv operands (tf32 operands are fp32 words)
        self.x_in = torch.empty(NB, H, W, 3, device=dev, dtype=f32)
        self.s_lr = torch.empty(NB, H, W, NUMK, device=dev, dtype=op)
        self.t1_lr = torch.empty_like(self.s_lr)
        self.t2_lr = torch.empty_like(self.s_lr)
        self.s_lr32 = torch.empty(NB, H, W, NUMK, device=dev, dtype=f32) if eng.stream_lr_fp32 else None
        self.out = torch.empty(NB, 4 * H, 4 * W, 3, device=dev, dtype=f32)   # [N,4H,4W,3] patch slots
        self.hr = {}             # (n, eh, ew) -> _HRStage reading this stage's stream
        lib = eng.lib
        names = [s[0] for s in layer_specs()]
        npix = NB * H * W
        if eng.tf32:
            self.steps.append(lambda st: L.check(lib.sr_head1x1_fwd(
                L.ptr(self.x_in), L.ptr(eng.head_w), L.ptr(eng.head_b), npix, None, L.ptr(self.s_lr32), st)))
            self.steps.append(lambda st: L.check(lib.sr_round_tf32(
                L.ptr(self.s_lr32), npix * NUMK, L.ptr(self.s_lr), st)))
        else:
            self.steps.append(lambda st: L.check(lib.sr_head1x1_fwd(
                L.ptr(self.x_in), L.ptr(eng.head_w), L.ptr(eng.head_b), npix, L.ptr(self.s_lr), L.ptr(self.s_lr32), st)))
        i, lr = 1, (NB, H, W)
        for b in range(16):
            self._block53(names, i, self.s_lr, self.s_lr32, self.t1_lr, self.t2_lr, lr, comp=exts[b])
            i += 4
        for b in range(16, 22):
            self._block_light(names, i, self.s_lr, self.s_lr32, self.t1_lr, lr, comp=exts[b])
            i += 2
        self.first_hr_layer = i

    def hr_stage(self, n, eh, ew):
        key = (n, eh, ew)
        st = self.hr.get(key)
        if st is None:
            if len(self.hr) >= 16:
                self.hr.pop(next(iter(self.hr)))
            try:
                st = _HRStage(self, n, eh, ew)
            except torch.cuda.OutOfMemoryError:
                for other in self.eng._graphs.values():     # free every other cached HR stage, keep this LR stage
                    other.hr.clear()
                torch.cuda.empty_cache()
                st = _HRStage(self, n, eh, ew)
            self.hr[key] = st
        return st


class _HRStage(_Stage):
    """High-resolution stage for n of the LR stage's patches, on the top-left eh x ew corner of their x4
    upsampling: bilinear x4 (TF1 legacy) -> 2 5/3 blocks -> tail conv3+relu, written into the LR stage's
    full-size output slots.  (eh, ew) = (4H, 4W) is the whole patch = model.predict; the tiled inference path
    uses 272x272: only [8,264) of a 384-pixel patch axis survives the stitch (img_utils.py:700-722) and the HR
    stage has a receptive-field radius of 7 pixels (2 blocks x (1+2) + 1), so nothing beyond 271 can reach a
    surviving pixel -- the surviving pixels are bit-identical to the full-patch computation."""

    def __init__(self, lrs, n, eh, ew):
        eng = lrs.eng
        self._init_stage(eng)
        dev, bf, f32 = eng.device, torch.bfloat16, torch.float32
        self.n, self.eh, self.ew = n, eh, ew
        self.src_index = torch.arange(n, device=dev, dtype=torch.int32)   # which LR patches (updated per run)
        self.idx_host = list(range(n))
        self.s_hr = torch.empty(n, eh, ew, NUMK, device=dev, dtype=f32 if eng.tf32 else bf)
        self.t1_hr = torch.empty_like(self.s_hr)
        self.t2_hr = torch.empty_like(self.s_hr)
        self.s_hr32 = torch.empty(n, eh, ew, NUMK, device=dev, dtype=f32) if eng.stream_hr_fp32 else None
        lib = eng.lib
        names = [s[0] for s in layer_specs()]
        src = lrs.s_lr32 if lrs.s_lr32 is not None else lrs.s_lr
        src_is_bf16 = 0 if lrs.s_lr32 is not None else 1
        H, W = lrs.H, lrs.W
        if eng.tf32:
            self.steps.append(lambda st: L.check(lib.sr_bilinear4_crop_fwd(
                L.ptr(src), 0, L.ptr(self.src_index), n, H, W, NUMK, eh, ew, None, L.ptr(self.s_hr32), st)))
            self.steps.append(lambda st: L.check(lib.sr_round_tf32(
                L.ptr(self.s_hr32), n * eh * ew * NUMK, L.ptr(self.s_hr), st)))
        else:
            self.steps.append(lambda st: L.check(lib.sr_bilinear4_crop_fwd(
                L.ptr(src), src_is_bf16, L.ptr(self.src_index), n, H, W, NUMK, eh, ew, L.ptr(self.s_hr),
                L.ptr(self.s_hr32), st)))
        i, hr = lrs.first_hr_layer, (n, eh, ew)
        # a cropped extent e (< 4H) was chosen as >= (last surviving pixel + 1) + 7, so the tail only has to be right on
        # [0, e-7), the second block on e-6 (its t1 / t2 on e-4 / e-5), the first block on e-3 (t1 / t2 on e-1 / e-2)
        def cut(k):
            return (eh - k if eh < 4 * H else eh, ew - k if ew < 4 * W else ew)
        comps = [(cut(3), cut(1), cut(2)), (cut(6), cut(4), cut(5))]
        for b in range(2):
            self._block53(names, i, self.s_hr, self.s_hr32, self.t1_hr, self.t2_hr, hr, comp=comps[b])
            i += 4
        self._conv([(names[i], self.s_hr)], hr, out_f32=lrs.out, relu=1, cout=3, out_index=self.src_index,
                   out_hw=(4 * H, 4 * W), comp=cut(7))


def lr_extents(need, full):
    """Compute extents (rows, cols) of every conv of the 22 LR blocks, given the region `need` of the final LR
    stream that is read afterwards.  Returns, per block, (out, t1, t2) for 5/3 blocks (0..15) and (out, t1) for light
    blocks (16..21).  Walking backwards: the fused conv of a 5/3 block reads t1 through a 5x5 (needs out+2) and t2
    through a 3x3 (out+1), which read the block input through a 3x3 / 5x5 (out+3); a light block needs out+1 / out+2."""
    def clip(e, add):
        return (min(full[0], e[0] + add), min(full[1], e[1] + add))
    e = (min(full[0], need[0]), min(full[1], need[1]))
    exts = [None] * 22
    for b in reversed(range(22)):
        if b >= 16:
            exts[b] = (e, clip(e, 1))
            e = clip(e, 2)
        else:
            exts[b] = (e, clip(e, 2), clip(e, 1))
            e = clip(e, 3)
    return exts


def hr_extent(tile, count, image_dim, patch=96, step=64, scale=4, radius=7, crop=8):
    """Rows (or columns) of a tile's x4 output that the HR stage must produce so that every pixel the tile
    contributes to the final [0, scale*image_dim) image is exact: (end of its owned span, clipped to the image) +
    receptive-field radius, rounded up to a multiple of 16.  272 for interior tiles of the 96/64/x4 geometry; the
    last live tile of an axis needs less when the image ends inside it, or the whole patch when it owns its end."""
    P, S = patch * scale, step * scale
    own_hi = S * (tile + 1) + crop if tile < count - 1 else S * tile + P       # exclusive, canvas coordinates
    own_hi = min(own_hi, scale * image_dim)
    need = max(own_hi - S * tile, 1) + radius
    return min(P, (need + 15) // 16 * 16)


class PackTable:
    """Device-side description of every (layer -> packed buffer) repack, so that all layers go in one launch."""

    def __init__(self, eng, packed, flip):
        import ctypes as C
        self.lib = eng.lib
        items, starts, total = [], [0], 0
        for name, k, cin, cout in eng.specs:
            if cin != NUMK:
                continue
            it = L.PackItem()
            it.hwio, it.dst = eng.master[name][0].data_ptr(), packed[name].data_ptr()
            it.ksize, it.cout, it.transpose_flip = k, cout, 1 if flip else 0
            items.append(it)
            total += packed[name].numel() // 2
            starts.append(total)
        arr = (L.PackItem * len(items))(*items)
        raw = np.frombuffer(memoryview(arr), dtype=np.uint8).copy()
        self.items = torch.from_numpy(raw).to(eng.device)
        self.starts = torch.tensor(starts, dtype=torch.int64).to(eng.device)
        self.n, self.total = len(items), total

    def run(self):
        L.check(self.lib.sr_pack_conv_weights_batched(L.ptr(self.items), L.ptr(self.starts), self.n, self.total,
                                                      L.stream_ptr()))


def plan_tiles(h, w, patch=96, step=64, scale=4, full_canvas=False):
    """Tile plan of one h x w image: (virtual canvas (H', W') handed to the gather, (cnt_h, cnt_w), per-tile HR
    extents in the gather's column-major order n = wi*cnt_h + hi, or None for full_canvas).

    full_canvas: the reference's canvas and tile grid (models.py:225-256, img_utils.py:622-648).  Otherwise only the
    tiles that own a pixel of the final [0,scale*h) x [0,scale*w) image (models.py:412): ownership switches from tile
    i-1 to tile i at scale*step*i + 8 (img_utils.py:700-722, later patches overwrite earlier ones), so the last live
    tile index per axis is floor((scale*dim - 1 - 8) / (scale*step)); the trailing tiles sit in the zero padding and
    are overwritten or cropped away."""
    from . import ops
    ch, cw = ops.canvas_size(h, w, patch, step)
    cnt_h, cnt_w = ops.patch_count(ch, patch, step), ops.patch_count(cw, patch, step)
    if full_canvas:
        return (ch, cw), (cnt_h, cnt_w), None
    S, c8 = step * scale, 8

    def live(dim, cnt):
        last = scale * dim - 1
        return 1 if last < S + c8 else min(cnt, (last - c8) // S + 1)

    lh, lw = live(h, cnt_h), live(w, cnt_w)
    gh, gw = (lh - 1) * step + patch + 1, (lw - 1) * step + patch + 1   # smallest canvas with exactly those counts
    ext = []
    for wi in range(lw):
        ew = hr_extent(wi, lw, w, patch, step, scale)
        for hi in range(lh):
            ext.append((hr_extent(hi, lh, h, patch, step, scale), ew))
    return (gh, gw), (lh, lw), ext


class Engine:
    """Device-resident DifvdsrDouble weights + cached per-shape graphs."""

    def __init__(self, weights=None, device=None, stream="lr32", a_mode=0, nacc=2, pair=1,
                 max_pixels=192 * 96 * 96, use_graphs=True, precision=None, sequencer=None):
        """sequencer: "c" (default) -- the launch sequence, its plans and CUDA graphs live in libsr100 behind
        sr_model_forward / sr_model_forward_backward (csrc/model.cu; one ctypes call per forward); "python"
        (SR100_PY_SEQUENCE=1) -- the same sequence issued launch by launch from this module (_LRStage / _HRStage),
        kept for per-launch tooling and as the bit-for-bit cross-check of the C entry points."""
        self.lib = L.require_device()
        sequencer = sequencer or ("python" if os.environ.get("SR100_PY_SEQUENCE", "0") == "1" else "c")
        if sequencer not in ("c", "python"):
            raise ValueError("sequencer must be 'c' or 'python', got %r" % (sequencer,))
        self.sequencer = sequencer
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        assert stream in ("bf16", "lr32", "fp32")
        precision = precision or os.environ.get("SR100_PRECISION", "bf16")
        if precision not in ("bf16", "tf32"):
            raise ValueError("precision must be 'bf16' or 'tf32', got %r" % (precision,))
        self.precision, self.tf32 = precision, precision == "tf32"
        if self.tf32:
            stream, a_mode, nacc = "fp32", 0, 2      # fp32 residual stream at both resolutions
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
        self._pack_table = None
        self._py_packed = self.sequencer != "c"     # C sequencer: the Python-side packed copy is built on first use
        self._init_bias_pairs()
        self._graphs = {}
        self.sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
        # second stream for independent launches of small inputs (_Stage._launch_all); SR100_NO_OVERLAP=1 disables it
        self.side_stream = None if os.environ.get("SR100_NO_OVERLAP", "0") == "1" else torch.cuda.Stream(self.device)
        self.model = C.c_void_p()            # sr_model*: owns packed weights, plans and graphs of the C sequencer
        self._ws = self._x_buf = self._out_buf = self._u8_buf = None
        self._stitch_cache = {}
        self.last_calls, self.last_stages = [], []
        self.set_weights_dict(weights if weights is not None else glorot_uniform_weights())

    def _ensure_model(self):
        if not self.model:
            cfg = L.ModelConfig()
            self.lib.sr_model_default_config(C.byref(cfg))
            cfg.precision = 1 if self.tf32 else 0
            cfg.stream_lr_fp32, cfg.stream_hr_fp32 = int(self.stream_lr_fp32), int(self.stream_hr_fp32)
            cfg.a_mode, cfg.nacc, cfg.pair = self.a_mode, self.nacc, self.pair
            cfg.use_graphs = int(self.use_graphs)
            cfg.overlap_heads = int(self.side_stream is not None)
            cfg.fused_colsum = int(os.environ.get("SR100_FUSED_COLSUM", "1") != "0")
            cfg.overlap_train = int(os.environ.get("SR100_OVERLAP_TRAIN", str(cfg.overlap_train)) != "0")
            cfg.chain_lr = int(os.environ.get("SR100_CHAIN_LR", "0") == "1")
            with torch.cuda.device(self.device):
                L.check(self.lib.sr_model_create(L.ptr(self.param_arena), C.byref(cfg), C.byref(self.model)))
        return self.model

    def __del__(self):
        try:
            if self.model:
                self.lib.sr_model_destroy(self.model)
                self.model = None
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def release(self):
        """Give the cached activation buffers back to the allocator (plans stay; they are rebuilt on demand)."""
        for st in self._graphs.values():
            st.hr.clear()
        self._graphs.clear()
        self._ws = self._x_buf = self._out_buf = self._u8_buf = None
        torch.cuda.empty_cache()

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

    def repack(self, c_model=True):
        """(Re)build the tensor-core weight layout from the fp32 masters (after load / optimizer step): one launch
        for all 85 layers (sr_pack_conv_weights_batched).  The C sequencer keeps its own packed copy (sr_model_refresh);
        the Python-side copy exists only once something asked for it (ensure_py_packed)."""
        if c_model and self.sequencer == "c":
            L.check(self.lib.sr_model_refresh(self._ensure_model(), L.stream_ptr()))
        if self.sequencer == "c" and not self._py_packed:
            return
        if self.tf32:
            st = L.stream_ptr()
            for name, k, cin, cout in self.specs:
                if cin != NUMK:
                    continue
                if name not in self.packed:
                    self.packed[name] = torch.empty(self.lib.sr_packed_weight_bytes_tf32(k, cout), dtype=torch.uint8,
                                                    device=self.device)
                L.check(self.lib.sr_pack_conv_weights_tf32(L.ptr(self.master[name][0]), k, cout,
                                                           L.ptr(self.packed[name]), st))
            self._refresh_bias_sums()
            return
        if self._pack_table is None:
            for name, k, cin, cout in self.specs:
                if cin == NUMK and name not in self.packed:
                    self.packed[name] = torch.empty(self.lib.sr_packed_weight_bytes(k, cout), dtype=torch.uint8,
                                                    device=self.device)
            self._pack_table = PackTable(self, self.packed, flip=False)
        self._pack_table.run()
        self._refresh_bias_sums()

    def ensure_py_packed(self):
        """Packed weights for plans created from Python (the "python" sequencer, Trainer's Python launch lists)."""
        if not self._py_packed:
            self._py_packed = True
            self.repack(c_model=False)

    def _init_bias_pairs(self):
        """The 18 two-source launches (fused tails of the 5/3 blocks) add two biases: one [18,128] buffer, refreshed
        by three launches after every weight change."""
        names = [s[0] for s in self.specs]
        pairs, i = [], 1
        for _ in range(16):
            pairs.append((names[i + 1], names[i + 3]))
            i += 4
        i += 12
        for _ in range(2):
            pairs.append((names[i + 1], names[i + 3]))
            i += 4
        ar = torch.arange(NUMK, device=self.device)
        self._pair_ia = torch.stack([self.param_slices[a][2] + ar for a, _ in pairs])
        self._pair_ib = torch.stack([self.param_slices[b][2] + ar for _, b in pairs])
        self._pair_buf = torch.zeros(len(pairs), NUMK, device=self.device, dtype=torch.float32)
        self._bias_cache = {p: self._pair_buf[j] for j, p in enumerate(pairs)}

    def _refresh_bias_sums(self):
        torch.add(self.param_arena[self._pair_ia], self.param_arena[self._pair_ib], out=self._pair_buf)

    def bias_for(self, names):
        """Device bias vector of a conv launch: the layer's own bias (a view into the parameter arena, always
        current), or for a two-source launch the sum of both biases (refreshed by repack)."""
        if len(names) == 1:
            return self.master[names[0]][1]
        return self._bias_cache[names]

    def get_weights_dict(self):
        return {n: (self.master[n][0].cpu().numpy(), self.master[n][1].cpu().numpy()) for n, _, _, _ in self.specs}

    # ---------------------------------------------------------------- forward
    def graph(self, NB, H, W, need=None):
        """The LR stage (buffers + plans) for NB patches of HxW; HR stages hang off it.  need: see _LRStage."""
        self.ensure_py_packed()
        need = (H, W) if need is None else (min(H, need[0]), min(W, need[1]))
        key = (NB, H, W) + need
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 4:  # bound device memory: keep the most recent shapes only
                self._graphs.pop(next(iter(self._graphs)))
            g = self._with_oom_retry(lambda: _LRStage(self, NB, H, W, need))
            self._graphs[key] = g
        return g

    def _with_oom_retry(self, make):
        """Stage buffers are cached per shape; if a new shape does not fit next to the cached ones, drop every cached
        stage (their buffers return to the allocator) and try once more."""
        try:
            return make()
        except torch.cuda.OutOfMemoryError:
            for st in self._graphs.values():
                st.hr.clear()
            self._graphs.clear()
            torch.cuda.empty_cache()
            return make()

    def sub_batch(self, H, W):
        return max(1, self.max_pixels // (H * W))

    def forward_device(self, x, out=None, extents=None):
        """x: device float32 [N,H,W,3] in [0,1] -> device float32 [N,4H,4W,3] (model.predict).
        extents: optional per-patch (eh, ew): only the top-left eh x ew corner of patch n's output is computed
        (the rest of its slot is unspecified) -- the tiled path's dead-region elimination, see _HRStage."""
        N, H, W, _ = x.shape
        if out is None:
            out = torch.empty(N, 4 * H, 4 * W, 3, device=self.device, dtype=torch.float32)
        nb = min(N, self.sub_batch(H, W))
        self.last_stages, self.last_calls = [], []
        if self.sequencer == "c":
            for i in range(0, N, nb):
                n = min(nb, N - i)
                self._forward_c(x[i:i + n], out[i:i + n], None if extents is None else extents[i:i + n])
            return out
        for i in range(0, N, nb):
            n = min(nb, N - i)
            if extents is None:
                groups = {(4 * H, 4 * W): list(range(n))}
                need = None
            else:
                groups = {}
                for j in range(n):
                    groups.setdefault(tuple(extents[i + j]), []).append(j)
                # the bilinear reads LR cells [0, e/4] of the stream: the LR layers shrink towards that region
                need = (max(e[0] for e in groups) // 4 + 1, max(e[1] for e in groups) // 4 + 1)
            g = self.graph(n, H, W, need)
            g.x_in.copy_(x[i:i + n])
            g.run()
            self.last_stages.append(g)
            for (eh, ew), idx in groups.items():
                hs = g.hr_stage(len(idx), eh, ew)
                if hs.idx_host != idx:          # steady workloads repeat the same tile classes: no host copy, no sync
                    hs.src_index.copy_(torch.tensor(idx, dtype=torch.int32))
                    hs.idx_host = list(idx)
                hs.run()
                self.last_stages.append(hs)
            out[i:i + n].copy_(g.out)
        return out

    # ---------------------------------------------------------------- the C sequencer (sr_model_forward)
    def _forward_desc(self, n, H, W, extents):
        """sr_forward_desc of one sub-batch on the engine's shared buffers (+ the ctypes arrays it points to)."""
        d = L.ForwardDesc()
        d.NB, d.H, d.W = n, H, W
        keep, groups = [], {}
        if extents is not None:
            for j in range(n):
                groups.setdefault((int(extents[j][0]), int(extents[j][1])), []).append(j)
            arr = lambda v: (C.c_int * len(v))(*v)
            eh, ew = arr([k[0] for k in groups]), arr([k[1] for k in groups])
            cnt, idx = arr([len(v) for v in groups.values()]), arr([j for v in groups.values() for j in v])
            d.n_groups, d.group_eh, d.group_ew, d.group_n, d.group_index = len(groups), eh, ew, cnt, idx
            keep = [eh, ew, cnt, idx]
        return d, keep, groups

    def _bind_buffers(self, d, need_out=True):
        """Point the descriptor at the shared grow-only buffers (workspace, input, output): stable pointers keep the
        library's plan / graph cache hot; a forward runs start to end on one stream, so sharing is safe."""
        need = self.lib.sr_model_forward_workspace_bytes(self._ensure_model(), C.byref(d))
        if need == 0:
            L.check(-1)
        nx = d.NB * d.H * d.W * 3

        def grow():
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            if self._x_buf is None or self._x_buf.numel() < nx:
                self._x_buf = None
                self._x_buf = torch.empty(nx, dtype=torch.float32, device=self.device)
            if need_out and (self._out_buf is None or self._out_buf.numel() < 16 * nx):
                self._out_buf = None
                self._out_buf = torch.empty(16 * nx, dtype=torch.float32, device=self.device)

        try:
            grow()
        except torch.cuda.OutOfMemoryError:
            self.release()
            grow()
        d.x = self._x_buf.data_ptr()
        d.out = self._out_buf.data_ptr() if need_out else None
        d.workspace, d.workspace_bytes = self._ws.data_ptr(), self._ws.numel()

    def _forward_c(self, x, out, extents, stitch=None):
        """One sub-batch through sr_model_forward.  stitch = (device pointer of this sub-batch's sr_stitch_tile rows,
        device pointer of the uint8 image buffer, mul): the tail convs write the owned uint8 pixels themselves and
        no fp32 patch tensor is produced (out is None)."""
        n, H, W, _ = x.shape
        d, keep, groups = self._forward_desc(n, H, W, extents)
        if stitch is not None:
            d.stitch_tiles, d.stitch_u8, d.stitch_mul = stitch
        self._bind_buffers(d, need_out=stitch is None)
        nx = n * H * W * 3
        self._x_buf[:nx].view(n, H, W, 3).copy_(x)
        L.check(self.lib.sr_model_forward(self.model, C.byref(d), L.stream_ptr()))
        if stitch is None:
            out.copy_(self._out_buf[:16 * nx].view(n, 4 * H, 4 * W, 3))
        self.last_calls.append((d, keep, groups))

    def forward_stitched(self, x, extents, tiles_dev, u8_buf, mul=255.0):
        """The conv stack over all patches x [N,H,W,3] with the quantise + stitch fused into the tail convs
        (sr_forward_desc.stitch_*): patch n writes the pixels stitch tile n owns into u8_buf (a uint8 tensor, or a raw
        device pointer -- e.g. rank 0's image mapped over NVLink, upscale_image_sharded)."""
        N, H, W, _ = x.shape
        nb = min(N, self.sub_batch(H, W))
        self.last_stages, self.last_calls = [], []
        tsz = C.sizeof(L.StitchTile)
        for i in range(0, N, nb):
            n = min(nb, N - i)
            self._forward_c(x[i:i + n], None, None if extents is None else extents[i:i + n],
                            stitch=(tiles_dev.data_ptr() + i * tsz,
                                    u8_buf if isinstance(u8_buf, int) else u8_buf.data_ptr(), float(mul)))

    @staticmethod
    def stitch_tiles_host(metas, patch, step, scale, crop=8):
        """sr_stitch_tile rows (numpy structured array) of all patches of a list of images, in gather order, and the
        byte size of the uint8 buffer that holds the images.  metas: [(out_h, out_w, (cnt_h, cnt_w), x_shift)] per
        image (x_shift: image column of canvas column 0, negative for a column strip of a sharded image).
        Ownership (img_utils.py:700-722, later patches overwrite earlier ones): patch i of an axis writes
        [c_i, P - c_i) of its P pixels (c_0 = 0, else 8) and loses everything from S + 8 on to patch i + 1."""
        P, S = patch * scale, step * scale
        dt = np.dtype([("img_offset", "<i8"), ("img_h", "<i4"), ("img_w", "<i4"), ("y0", "<i4"), ("x0", "<i4"),
                       ("oy0", "<i4"), ("oy1", "<i4"), ("ox0", "<i4"), ("ox1", "<i4")])
        assert dt.itemsize == C.sizeof(L.StitchTile)

        def own(i, cnt):
            lo = np.where(i == 0, 0, crop)
            end = np.where(i == 0, P, P - crop)
            return lo, np.where(i < cnt - 1, np.minimum(end, S + crop), end)

        rows, off, offsets = [], 0, []
        for oh, ow, (cnt_h, cnt_w), x_shift in metas:
            n = np.arange(cnt_h * cnt_w)
            wi, hi = n // cnt_h, n % cnt_h
            t = np.zeros(n.size, dtype=dt)
            t["img_offset"], t["img_h"], t["img_w"] = off, oh, ow
            t["y0"], t["x0"] = S * hi, S * wi + x_shift
            t["oy0"], t["oy1"] = own(hi, cnt_h)
            t["ox0"], t["ox1"] = own(wi, cnt_w)
            rows.append(t)
            offsets.append(off)
            off += (oh * ow * 3 + 255) // 256 * 256
        return np.concatenate(rows), offsets, off

    def timed_launches(self):
        """[(ms, flops)] of every launch of the most recent forward_device call, re-run launch by launch between
        CUDA events (sr_model_forward_timed; flops = 0 for launches that are not tensor-core convs)."""
        res = []
        if self.sequencer != "c":
            st = L.stream_ptr()
            evs = []
            for stg in self.last_stages:
                for step in stg.steps:
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    step(st)
                    b.record()
                    owner = getattr(step, "__self__", None)
                    evs.append((a, b, owner.flops if isinstance(owner, _Plan) else 0.0))
            torch.cuda.synchronize()
            return [(a.elapsed_time(b), f) for a, b, f in evs]
        for d, keep, _ in self.last_calls:
            info = L.ModelRunInfo()
            L.check(self.lib.sr_model_forward_info(self.model, C.byref(d), C.byref(info)))
            ms, fl, cnt = (C.c_float * info.launches)(), (C.c_double * info.launches)(), C.c_int(0)
            L.check(self.lib.sr_model_forward_timed(self.model, C.byref(d), L.stream_ptr(), ms, fl, info.launches,
                                                    C.byref(cnt)))
            res.extend(zip(list(ms), list(fl)))
        return res

    def last_run_summary(self):
        """(patches run, sorted HR-stage extents) of the most recent forward_device call."""
        if self.sequencer == "c":
            tiles = sum(d.NB for d, _, _ in self.last_calls)
            ext = sorted({k for _, _, g in self.last_calls for k in g} or
                         {(4 * d.H, 4 * d.W) for d, _, _ in self.last_calls})
            return tiles, ext
        tiles = sum(st_.NB for st_ in self.last_stages if hasattr(st_, "x_in"))
        return tiles, sorted({(st_.eh, st_.ew) for st_ in self.last_stages if hasattr(st_, "eh")})

    def graph_ready(self):
        """True when the next forward of the most recent shape replays a captured CUDA graph."""
        if self.sequencer == "c":
            ok = bool(self.last_calls)
            for d, _, _ in self.last_calls:
                info = L.ModelRunInfo()
                L.check(self.lib.sr_model_forward_info(self.model, C.byref(d), C.byref(info)))
                ok = ok and bool(info.graph_replay)
            return ok
        return bool(self.last_stages) and all(st.cuda_graph is not None for st in self.last_stages)

    def upscale_images_device(self, imgs_u8, patch=96, step=64, scale=4, full_canvas=False):
        """The device part of upscaleStepPatch (models.py:225-391) for a list of uint8 [h,w,3] device images:
        zero-padded canvas -> 96/64 patch gather (/255) -> conv stack over ALL tiles of all images -> x255,
        stitch with the 8-px crop, clip -> uint8.

        full_canvas=True reproduces every tile and returns the uncropped uint8 canvases (what
        upscaleStepPatch(return_image=True) hands back, models.py:405-407).  The default returns the final
        [4h,4w,3] images (models.py:412) and eliminates work that cannot reach them: tiles whose owned span lies
        entirely in the zero padding are not run, and the HR stage runs on the 272x272 corner of each patch that
        the stitch can see (hr_extent).  Both give bit-identical pixels inside the final image."""
        from . import ops
        if len(imgs_u8) == 0:
            return []
        metas, extents, total = [], [], 0
        for img in imgs_u8:
            h, w, _ = img.shape
            (gh, gw), counts, ext = plan_tiles(h, w, patch, step, scale, full_canvas)
            ch, cw = ops.canvas_size(h, w, patch, step)
            n = counts[0] * counts[1]
            metas.append((h, w, ch, cw, counts, n, (gh, gw)))
            total += n
            if not full_canvas:
                extents.extend(ext)
        # gather straight into one patch tensor; runs of same-shaped images (a batch, BASELINE config 3) go in one launch
        allp = torch.empty(total, patch, patch, 3, device=self.device, dtype=torch.float32)
        i, off = 0, 0
        while i < len(imgs_u8):
            j = i + 1
            while j < len(imgs_u8) and imgs_u8[j].shape == imgs_u8[i].shape:
                j += 1
            n, g = metas[i][5], metas[i][6]
            batch = imgs_u8[i].unsqueeze(0) if j - i == 1 else torch.stack(imgs_u8[i:j])
            _, got = ops.patch_gather_u8_batched(batch.contiguous(), g, (patch, patch), step, divisor=255.0,
                                                 out=allp[off:off + (j - i) * n])
            assert got == metas[i][4]
            off += (j - i) * n
            i = j
        if self.sequencer == "c" and os.environ.get("SR100_FUSED_STITCH", "1") != "0":
            # x255, clip -> uint8 and the 8-px-crop stitch happen in the tail convs' epilogues: every patch writes
            # the pixels it owns straight into its image (no fp32 patch tensor, no stitch pass)
            key = (tuple((m[0], m[1]) for m in metas), patch, step, scale, bool(full_canvas))
            cached = self._stitch_cache.get(key)
            if cached is None:
                sm = [((scale * ch, scale * cw) if full_canvas else (scale * h, scale * w)) + (counts, 0)
                      for h, w, ch, cw, counts, n, _ in metas]
                rows, offsets, nbytes = self.stitch_tiles_host(sm, patch, step, scale)
                if len(self._stitch_cache) >= 8:
                    self._stitch_cache.pop(next(iter(self._stitch_cache)))
                cached = (torch.from_numpy(rows.view(np.uint8)).to(self.device), offsets, nbytes,
                          [(m[0], m[1]) for m in sm])
                self._stitch_cache[key] = cached
            tiles_dev, offsets, nbytes, dims = cached
            if self._u8_buf is None or self._u8_buf.numel() < nbytes:
                self._u8_buf = None
                self._u8_buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._u8_buf[:nbytes].zero_()          # pixels no patch owns stay 0 (the reference's zero canvas)
            self.forward_stitched(allp, None if full_canvas else extents, tiles_dev, self._u8_buf)
            return [self._u8_buf[o:o + oh * ow * 3].view(oh, ow, 3).clone() for o, (oh, ow) in zip(offsets, dims)]
        out = self.forward_device(allp, extents=None if full_canvas else extents)
        res, off = [], 0
        for h, w, ch, cw, counts, n, _ in metas:
            oh, ow = (ch, cw) if full_canvas else (h, w)
            _, u8 = ops.patch_stitch(out[off:off + n], counts, (patch, patch), step, scale, (oh, ow), mul=255.0,
                                     want_f32=False, want_u8=True)
            res.append(u8)
            off += n
        return res

    def shard_plan(self, h, w, world, patch=96, step=64, scale=4):
        """[(tile_lo, tile_hi, x0, x1)] per rank for one h x w image, + the tile plan: contiguous ranges of the
        column-major live-tile index and the output columns each range can own."""
        from . import ops
        from .dist import shard_range
        (gh, gw), counts, ext = plan_tiles(h, w, patch, step, scale, False)
        n_tiles = counts[0] * counts[1]
        shards = []
        for r in range(world):
            lo, hi = shard_range(n_tiles, r, world)
            x0, x1 = ops.shard_strip(counts, (patch, patch), step, scale, scale * w, lo, hi) if hi > lo else (0, 0)
            shards.append((lo, hi, x0, x1))
        return (gh, gw), counts, ext, shards

    def stitch_shard(self, out, counts, h, lo, hi, x0, width, patch=96, step=64, scale=4):
        """uint8 strip [scale*h, width, 3] of the pixels tiles [lo, hi) own (zeros elsewhere); out = their patches."""
        from . import ops
        strip = torch.zeros(scale * h, width, 3, device=self.device, dtype=torch.uint8)
        if hi > lo:
            ops.patch_stitch_range(out, counts, (patch, patch), step, scale, h, lo, hi, x0, width, mul=255.0, out=strip)
        return strip

    def upscale_image_sharded(self, img_u8, patch=96, step=64, scale=4, world=None, rank=None, group_gather=None):
        """One (large) image with its tiles sharded over the ranks of the process group (BASELINE config
        5): every rank runs a contiguous range of the column-major live-tile index through the conv stack; the tail
        convs' epilogues write the uint8 pixels those tiles OWN (3 bytes per owned pixel, not 12-byte fp32 patches
        with their 2.25x overlap).  Default under NCCL ('p2p'): straight into rank 0's image, mapped into every
        rank over NVLink peer memory (CUDA IPC) -- the transfer rides the tail convs' stores, one stream-ordered peer
        barrier ends the step, there is no gather and no merge pass.  Otherwise (SR100_SHARD_GATHER=nccl, ranks on
        several hosts): into a per-rank column strip, gathered on rank 0 and OR-ed into the image.  Ownership is a
        partition (and non-owned strip pixels are 0), so both are bit-identical to the single-rank stitch.  Returns
        the uint8 [4h,4w,3] device image on rank 0, None elsewhere.  (world / rank / group_gather: a logical split
        inside one process, for tests.)"""
        import torch.distributed as tdist
        from . import ops
        h, w, _ = img_u8.shape
        dist_on = tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1
        if world is None:
            world, rank = (tdist.get_world_size(), tdist.get_rank()) if dist_on else (1, 0)
        (gh, gw), counts, ext, shards = self.shard_plan(h, w, world, patch, step, scale)
        p, got = ops.patch_gather_u8(img_u8, (gh, gw), (patch, patch), step, divisor=255.0)
        assert got == counts
        fused = self.sequencer == "c" and os.environ.get("SR100_FUSED_STITCH", "1") != "0"

        def run_fused(lo, hi, x0, width):
            """Tiles [lo, hi) with the stitch fused into the tail convs: their owned pixels land in a zeroed uint8
            strip of output columns [x0, x0 + width)."""
            key = ("shard", h, w, patch, step, scale, lo, hi, x0, width)
            cached = self._stitch_cache.get(key)
            if cached is None:
                rows, _, _ = self.stitch_tiles_host([(scale * h, width, counts, -x0)], patch, step, scale)
                if len(self._stitch_cache) >= 8:
                    self._stitch_cache.pop(next(iter(self._stitch_cache)))
                cached = torch.from_numpy(np.ascontiguousarray(rows[lo:hi]).view(np.uint8)).to(self.device)
                self._stitch_cache[key] = cached
            nbytes = scale * h * width * 3
            if self._u8_buf is None or self._u8_buf.numel() < nbytes:
                self._u8_buf = None
                self._u8_buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._u8_buf[:nbytes].zero_()
            if hi > lo:
                self.forward_stitched(p[lo:hi], ext[lo:hi], cached, self._u8_buf)
            return self._u8_buf[:nbytes].view(scale * h, width, 3)

        if world == 1:
            if fused:
                return run_fused(0, counts[0] * counts[1], 0, scale * w).clone()
            out = self.forward_device(p, extents=ext)
            _, u8 = ops.patch_stitch(out, counts, (patch, patch), step, scale, (h, w), mul=255.0, want_f32=False,
                                     want_u8=True)
            return u8
        lo, hi, x0, x1 = shards[rank]
        if fused and group_gather is None and dist_on and self._peer_canvas_mode() == "p2p":
            st = self._peer_canvas(scale * h * scale * w * 3)
            if st is not None:
                # every rank's tail convs store the pixels its tiles own straight into rank 0's image over NVLink
                # (ownership is a partition: no two ranks write the same byte); one barrier says "all written",
                # rank 0 copies the image out, a second barrier hands the canvas back for the next call
                ptr, canvas, bar = st
                key = ("p2p", h, w, patch, step, scale, lo, hi)
                cached = self._stitch_cache.get(key)
                if cached is None:
                    rows, _, _ = self.stitch_tiles_host([(scale * h, scale * w, counts, 0)], patch, step, scale)
                    if len(self._stitch_cache) >= 8:
                        self._stitch_cache.pop(next(iter(self._stitch_cache)))
                    cached = torch.from_numpy(np.ascontiguousarray(rows[lo:hi]).view(np.uint8)).to(self.device)
                    self._stitch_cache[key] = cached
                if hi > lo:
                    self.forward_stitched(p[lo:hi], ext[lo:hi], cached, ptr)
                bar.arrive_wait()
                out = canvas[:scale * h * scale * w * 3].view(scale * h, scale * w, 3).clone() if rank == 0 else None
                bar.arrive_wait()
                return out
        wmax = max(s[3] - s[2] for s in shards)
        if fused:
            send = run_fused(lo, hi, x0, wmax)
        else:
            out = self.forward_device(p[lo:hi], extents=ext[lo:hi]) if hi > lo else p[:0]
            send = self.stitch_shard(out, counts, h, lo, hi, x0, wmax, patch, step, scale)
        if group_gather is not None:
            recv = group_gather(send)                      # tests: the logical ranks' strips, in rank order
        else:
            recv = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
            tdist.gather(send, recv, dst=0)
        if rank != 0 or recv is None:
            return None
        u8 = torch.zeros(scale * h, scale * w, 3, device=self.device, dtype=torch.uint8)
        for (rlo, rhi, rx0, rx1), strip in zip(shards, recv):
            if rhi > rlo:
                u8[:, rx0:rx1].bitwise_or_(strip[:, :rx1 - rx0])
        return u8

    @staticmethod
    def _peer_canvas_mode():
        import torch.distributed as tdist
        return os.environ.get("SR100_SHARD_GATHER") or ("p2p" if tdist.get_backend() == "nccl" else "nccl")

    def _peer_canvas(self, nbytes):
        """(canvas pointer valid in this process, rank 0's canvas tensor, PeerBarrier) -- rank 0's uint8 image mapped
        into every rank with CUDA IPC (collective on first use and when a larger image arrives); None if the ranks do
        not share a node or the mapping fails (the caller falls back to the strip gather, on every rank alike)."""
        st = getattr(self, "_peer_canvas_state", None)
        if st is not None and (st == "unavailable" or st[0] >= nbytes):
            return None if st == "unavailable" else st[1:]
        from . import peer
        try:
            ptr, canvas, bar = peer.connect_canvas(self.lib, self.device, nbytes)
            self._peer_canvas_state = (nbytes, ptr, canvas, bar)
            return ptr, canvas, bar
        except RuntimeError as e:
            import warnings
            warnings.warn("sr100: %s; tile-sharded images use the strip gather" % e, RuntimeWarning)
            self._peer_canvas_state = "unavailable"
            return None

    @property
    def sharded_gather_description(self):
        import torch.distributed as tdist
        dist_on = tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1
        st = getattr(self, "_peer_canvas_state", None)
        if dist_on and st is not None and st != "unavailable":
            return ("every rank's tail convs store the uint8 pixels its tiles own straight into rank 0's image over "
                    "NVLink peer memory (CUDA IPC), one stream-ordered peer barrier, no gather / merge pass")
        return ("every rank stitches the pixels its tiles own into a uint8 column strip, the strips are gathered on "
                "rank 0 (NCCL) and OR-ed into the image")

    def last_flops(self):
        """Algorithmic FLOPs (2*MAC) of the tensor-core launches of the most recent forward_device call."""
        if self.sequencer == "c":
            tot = 0.0
            for d, _, _ in self.last_calls:
                info = L.ModelRunInfo()
                L.check(self.lib.sr_model_forward_info(self.model, C.byref(d), C.byref(info)))
                tot += info.conv_flops
            return tot
        return float(sum(st.conv_flops for st in getattr(self, "last_stages", [])))

    def conv_flops(self, N, H, W):
        """Algorithmic FLOPs (2*MAC) of one forward over N patches of HxW (tensor-core convs + head)."""
        lr = N * H * W
        per_lr = 16 * (2 * 9 + 2 * 25) * NUMK * NUMK * 2 + 6 * 2 * 9 * NUMK * NUMK * 2 + 2 * 3 * NUMK
        per_hr = 2 * (2 * 9 + 2 * 25) * NUMK * NUMK * 2 + 9 * NUMK * 3 * 2
        return float(lr) * per_lr + float(lr) * 16 * per_hr
