"""Training step of DifvdsrDouble on sm_100a: forward with saved activations, MSE, backward (dgrad / wgrad on
the tensor cores), data-parallel gradient all-reduce, fused Keras-Adam.

Reference: BaseSuperResolutionModel.fit (models.py:131-157) = Keras fit_generator -> train_on_batch on the graph of
models.py:1159-1270 compiled with loss='mse', Adam(lr=1e-4, beta_1=0.9) (models.py:1212-1213; Keras-2 defaults
beta_2=0.999, epsilon=1e-7).

Backward of a 5/3 block  y = 0.1*(conv5_b(t1) + conv3_d(t2)) + 0.9*x,  t1 = relu(conv3_a(x)), t2 = relu(conv5_c(x)):
    g_t1 = 0.1 * dgrad_b(g_y) * (t1 > 0)          g_t2 = 0.1 * dgrad_d(g_y) * (t2 > 0)
    g_x  = dgrad_a(g_t1) + dgrad_c(g_t2) + 0.9 * g_y                       (one fused two-source launch)
    dW_b = 0.1 * wgrad(t1, g_y)   dW_d = 0.1 * wgrad(t2, g_y)   dW_a = wgrad(x, g_t1)   dW_c = wgrad(x, g_t2)
dgrad is the forward conv kernel with 180-degree-rotated, cin<->cout-transposed weights and a fused ReLU mask;
wgrad is csrc/wgrad_tc.cu.  Operands are bf16, accumulation fp32, master weights / gradients / Adam state fp32.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L
from .engine import NUMK, _Plan, layer_specs


class _WgradPlan:
    def __init__(self, lib, x, g, shape, ksize, scale, dw, workspace):
        self.lib = lib
        d = L.WgradDesc()
        d.x_bf16, d.g_bf16 = x.data_ptr(), g.data_ptr()
        d.NB, d.H, d.W = shape
        d.ksize, d.scale, d.accumulate = ksize, scale, 0
        d.dw_hwio, d.workspace, d.workspace_bytes = dw.data_ptr(), workspace.data_ptr(), workspace.numel()
        self.handle = C.c_void_p()
        L.check(lib.sr_wgrad_plan_create(C.byref(d), C.byref(self.handle)))
        info = L.WgradPlanInfo()
        L.check(lib.sr_wgrad_plan_info(self.handle, C.byref(info)))
        self.flops = info.flops

    def run(self, stream):
        L.check(self.lib.sr_wgrad_plan_run(self.handle, stream))

    def __del__(self):
        try:
            if self.handle:
                self.lib.sr_wgrad_plan_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001
            pass


class _Op:
    """A named launch (callable(stream)) so tools/bench_train.py can attribute step time per kernel kind."""

    def __init__(self, name, fn):
        self.name, self.fn = name, fn

    def __call__(self, st):
        return self.fn(st)


class _TrainGraph:
    """Buffers and launch lists (forward with saved activations, backward) for one (NB, H, W) minibatch shape."""

    def __init__(self, tr, NB, H, W):
        eng, lib, dev = tr.engine, tr.engine.lib, tr.engine.device
        bf, f32 = torch.bfloat16, torch.float32
        self.NB, self.H, self.W = NB, H, W
        HH, WW = 4 * H, 4 * W
        lr, hr = (NB, H, W), (NB, HH, WW)
        names = [s[0] for s in layer_specs()]

        def act(shape, dtype=bf, ch=NUMK):
            return torch.empty(*shape, ch, device=dev, dtype=dtype)

        self.x_in = act(lr, f32, 3)
        self.y_true = act(hr, f32, 3)
        self.out = act(hr, f32, 3)
        # saved forward tensors (bf16 = exactly what the convs consumed)
        S = [act(lr) for _ in range(23)]            # S[0] = head output, S[i+1] = output of LR block i
        T1 = [act(lr) for _ in range(22)]
        T2 = [act(lr) for _ in range(16)]
        s32 = act(lr, f32)                          # fp32 residual stream (rotating, forward only)
        SH = [act(hr) for _ in range(3)]
        TH1 = [act(hr) for _ in range(2)]
        TH2 = [act(hr) for _ in range(2)]
        # gradients
        gs, gs32 = act(lr), act(lr, f32)
        gt1, gt2 = act(lr), act(lr)
        gsh, gsh32 = act(hr), act(hr, f32)
        gth1, gth2 = act(hr), act(hr)
        self.loss_sum = torch.zeros(1, device=dev, dtype=torch.float64)
        self.keep = [S, T1, T2, s32, SH, TH1, TH2, gs, gs32, gt1, gt2, gsh, gsh32, gth1, gth2]
        self.fwd, self.bwd = [], []
        self.fwd_flops = self.bwd_flops = 0.0
        self.cuda_graph, self.ran_eager = None, False

        def conv(dst, srcs, shape, flip=False, out_bf16=None, out_f32=None, relu=0, alpha=1.0, beta=0.0,
                 res32=None, res16=None, mask=None, cout=NUMK, bias=True, colsum=None):
            """colsum = (bias-gradient view, scale): the launch also accumulates scale * column sums of its bf16 output
            (sr_conv_desc.colsum_f32) -- the bias gradient of the layer whose output gradient it writes."""
            d = L.ConvDesc()
            d.nsrc = len(srcs)
            for s, (name, x) in enumerate(srcs):
                d.in_[s] = x.data_ptr()
                d.wpacked[s] = (tr.packed_t if flip else eng.packed)[name].data_ptr()
                d.ksize[s] = eng.ksize[name]
            d.NB, d.H, d.W = shape
            d.cin, d.cout = NUMK, cout
            d.bias = eng.bias_for(tuple(n for n, _ in srcs)).data_ptr() if bias else None
            d.alpha, d.beta, d.relu = alpha, beta, relu
            d.res_f32 = res32.data_ptr() if res32 is not None else None
            d.res_bf16 = res16.data_ptr() if (res16 is not None and res32 is None) else None
            d.out_bf16 = out_bf16.data_ptr() if out_bf16 is not None else None
            d.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
            d.relu_mask_bf16 = mask.data_ptr() if mask is not None else None
            d.a_mode, d.nacc, d.pair = eng.a_mode, eng.nacc, eng.pair
            if colsum is not None:
                d.colsum_f32, d.colsum_scale = colsum[0].data_ptr(), colsum[1]
            p = _Plan(lib, d)
            if dst is self.fwd:
                self.fwd_flops += p.flops
            else:
                self.bwd_flops += p.flops
            kind = ("dgrad" if flip else "conv") + "+".join("k%d" % eng.ksize[n] for n, _ in srcs)
            dst.append(_Op("%s_%s" % (kind, "hr" if shape[1] != H else "lr"), p.run))

        def wgrad(x, g, shape, name, scale):
            p = _WgradPlan(lib, x, g, shape, eng.ksize[name], scale, tr.grad_w(name), tr.workspace)
            self.bwd_flops += p.flops
            self.bwd.append(_Op("wgrad_k%d_%s" % (eng.ksize[name], "hr" if shape[1] != H else "lr"), p.run))

        def colsum(g, shape, name, scale):
            npix = shape[0] * shape[1] * shape[2]
            db = tr.grad_b(name)
            self.bwd.append(_Op("colsum", lambda st: L.check(lib.sr_colsum_bf16(L.ptr(g), npix, scale, L.ptr(db), st))))

        # ------------------------------------------------------------------ forward
        npix = NB * H * W
        self.fwd.append(lambda st: L.check(lib.sr_head1x1_fwd(
            L.ptr(self.x_in), L.ptr(eng.head_w), L.ptr(eng.head_b), npix, L.ptr(S[0]), L.ptr(s32), st)))
        blocks = []  # (kind, first layer index, x, t1, t2, y, shape, f32 stream)
        i = 1
        for b in range(16):
            blocks.append(("53", i, S[b], T1[b], T2[b], S[b + 1], lr, True))
            i += 4
        for b in range(6):
            blocks.append(("light", i, S[16 + b], T1[16 + b], None, S[17 + b], lr, True))
            i += 2
        for b in range(2):
            blocks.append(("53", i, SH[b], TH1[b], TH2[b], SH[b + 1], hr, False))
            i += 4
        tail = names[i]

        def fwd_block(kind, i, x, t1, t2, y, shape, f32s):
            r32 = s32 if f32s else None
            if kind == "53":
                conv(self.fwd, [(names[i], x)], shape, out_bf16=t1, relu=1)
                conv(self.fwd, [(names[i + 2], x)], shape, out_bf16=t2, relu=1)
                conv(self.fwd, [(names[i + 1], t1), (names[i + 3], t2)], shape, out_bf16=y, out_f32=r32, alpha=0.1,
                     beta=0.9, res32=r32, res16=x)
            else:
                conv(self.fwd, [(names[i], x)], shape, out_bf16=t1, relu=1)
                conv(self.fwd, [(names[i + 1], t1)], shape, out_bf16=y, out_f32=r32, alpha=0.1, beta=1.0, res32=r32,
                     res16=x)

        for blk in blocks[:22]:
            fwd_block(*blk)
        self.fwd.append(lambda st: L.check(lib.sr_bilinear4_fwd(L.ptr(s32), 0, NB, H, W, NUMK, L.ptr(SH[0]), None, st)))
        for blk in blocks[22:]:
            fwd_block(*blk)
        conv(self.fwd, [(tail, SH[2])], hr, out_f32=self.out, relu=1, cout=3)

        # ------------------------------------------------------------------ backward
        npix_hr = NB * HH * WW
        self.n_local = npix_hr * 3
        # tail conv (128 -> 3, 3x3): the loss gradient is written as the im2col of the tail's backward (27 channels of a
        # 128-channel bf16 tensor), which turns both tail gradients into 1x1 problems for the tensor-core kernels
        tail_w, tail_b = tr.grad_w(tail), tr.grad_b(tail)       # views of the gradient arena (zeroed every step)
        self.bwd.append(_Op("tail_grad", lambda st: L.check(lib.sr_mse_tail_grad_col(
            L.ptr(self.out), L.ptr(self.y_true), NB, HH, WW, self.n_local, L.ptr(gth1), L.ptr(self.loss_sum),
            L.ptr(tail_b), st))))
        p = _WgradPlan(lib, SH[2], gth1, hr, 1, 1.0, tr.tail_d128, tr.workspace)      # D[ci][j], j = (ky,kx,co)
        self.bwd_flops += 2.0 * npix_hr * 27 * NUMK
        self.bwd.append(_Op("wgrad_tail", p.run))
        self.bwd.append(lambda st: tail_w.copy_(tr.tail_d128.view(NUMK, NUMK)[:, :27].reshape(NUMK, 3, 3, 3)
                                                .permute(1, 2, 0, 3)))
        # dgrad: 1x1 conv with B[j][ci] = W[ky][kx][ci][co] (packed at every weight refresh, Trainer.repack_t)
        d = L.ConvDesc()
        d.nsrc = 1
        d.in_[0], d.wpacked[0], d.ksize[0] = gth1.data_ptr(), tr.tail_colw_packed.data_ptr(), 1
        d.NB, d.H, d.W, d.cin, d.cout = NB, HH, WW, NUMK, NUMK
        d.alpha, d.beta, d.relu = 1.0, 0.0, 0
        d.out_bf16 = gsh.data_ptr()
        d.a_mode, d.nacc, d.pair = eng.a_mode, eng.nacc, eng.pair
        pl = _Plan(lib, d)
        self.bwd_flops += 2.0 * npix_hr * 27 * NUMK
        self.bwd.append(_Op("dgrad_tail", pl.run))

        fuse_cs = eng.nacc == 2 and eng.a_mode == 0 and os.environ.get("SR100_FUSED_COLSUM", "1") != "0"

        def g_bias(blk):
            """(bias-gradient view, scale, twin) of the layer(s) whose output gradient is the block's incoming g:
            the 0.1-scaled block tail -- both tail convs of a 5/3 block share it (twin = the second one)."""
            kind, i = blk[0], blk[1]
            if kind == "53":
                return tr.grad_b(names[i + 1]), 0.1, tr.grad_b(names[i + 3])
            return tr.grad_b(names[i + 1]), 0.1, None

        def bwd_block(blk, g_summed, nxt, last_hr=False):
            """blk's backward.  g_summed: the bias gradient from the incoming g was already accumulated by the launch
            that produced g.  nxt: the block processed next (it consumes the g this block writes): its tail bias
            gradient rides this block's last input-gradient launch."""
            kind, i, x, t1, t2, y, shape, f32s = blk
            g, g32 = (gs, gs32) if f32s else (gsh, None)
            a1, a2 = (gt1, gt2) if f32s else (gth1, gth2)
            o32 = g32 if f32s else (gsh32 if last_hr else None)
            db_g, sc_g, twin = g_bias(blk)
            # the launch that writes the next block's g can carry its column sums only if g stays bf16-resident in the
            # same buffer (not across the HR -> LR boundary, where g goes through the bilinear adjoint)
            cs_next = None
            if fuse_cs and nxt is not None and nxt[7] == f32s:
                cs_next = g_bias(nxt)[:2]
            if kind == "53":
                na, nb, nc, nd = names[i], names[i + 1], names[i + 2], names[i + 3]
                conv(self.bwd, [(nb, g)], shape, flip=True, out_bf16=a1, alpha=0.1, mask=t1, bias=False,
                     colsum=(tr.grad_b(na), 1.0) if fuse_cs else None)
                conv(self.bwd, [(nd, g)], shape, flip=True, out_bf16=a2, alpha=0.1, mask=t2, bias=False,
                     colsum=(tr.grad_b(nc), 1.0) if fuse_cs else None)
                wgrad(t1, g, shape, nb, 0.1)
                wgrad(t2, g, shape, nd, 0.1)
                if not g_summed:
                    colsum(g, shape, nb, 0.1)
                self.bwd.append(lambda st: twin.copy_(db_g))
                conv(self.bwd, [(na, a1), (nc, a2)], shape, flip=True, out_bf16=g, out_f32=o32, alpha=1.0, beta=0.9,
                     res32=g32, res16=g, bias=False, colsum=cs_next)
                wgrad(x, a1, shape, na, 1.0)
                wgrad(x, a2, shape, nc, 1.0)
                if not fuse_cs:
                    colsum(a1, shape, na, 1.0)
                    colsum(a2, shape, nc, 1.0)
            else:
                na, nb = names[i], names[i + 1]
                conv(self.bwd, [(nb, g)], shape, flip=True, out_bf16=a1, alpha=0.1, mask=t1, bias=False,
                     colsum=(tr.grad_b(na), 1.0) if fuse_cs else None)
                wgrad(t1, g, shape, nb, 0.1)
                if not g_summed:
                    colsum(g, shape, nb, 0.1)
                conv(self.bwd, [(na, a1)], shape, flip=True, out_bf16=g, out_f32=o32, alpha=1.0, beta=1.0, res32=g32,
                     res16=g, bias=False, colsum=cs_next)
                wgrad(x, a1, shape, na, 1.0)
                if not fuse_cs:
                    colsum(a1, shape, na, 1.0)
            return cs_next is not None

        # Gradient buckets for the overlapped all-reduce (Trainer.step_device): the arena is in layer order and the
        # backward runs last layer first, so "everything from layer L to the end of the previous bucket" is final
        # once the backward of the block that starts at L has run.  marks: (len(self.bwd) at that point, arena lo, hi).
        off = lambda bi: eng.param_slices[names[blocks[bi][1]]][0]
        self.marks = []
        done = bwd_block(blocks[23], False, blocks[22])
        bwd_block(blocks[22], done, None, last_hr=True)
        self.bwd.append(lambda st: L.check(lib.sr_bilinear4_bwd(L.ptr(gsh32), NB, H, W, NUMK, L.ptr(gs32), st)))
        self.bwd.append(lambda st: L.check(lib.sr_cast_f32_to_bf16(L.ptr(gs32), npix * NUMK, L.ptr(gs), st)))
        self.marks.append((len(self.bwd), off(22), eng.n_params))            # HR stage + tail: 9 layers, 9.4 MB
        done = False
        for bi in reversed(range(22)):
            done = bwd_block(blocks[bi], done, blocks[bi - 1] if bi > 0 else None)
            if bi == 11:
                # the launch that wrote block 10's tail bias gradient (cs_next) ran inside this block: that range
                # belongs to the LAST bucket, which is reduced after everything
                self.marks.append((len(self.bwd), off(11), off(22)))          # LR blocks 11..21: 39 MB
        hw, hb = tr.grad_w("level1"), tr.grad_b("level1")
        self.bwd.append(lambda st: L.check(lib.sr_head1x1_bwd(
            L.ptr(self.x_in), L.ptr(S[0]), L.ptr(gs32), None, npix, L.ptr(hw), L.ptr(hb), st)))
        self.marks.append((len(self.bwd), 0, off(11)))                        # head + LR blocks 0..10: 39 MB
        self.seg_graphs, self.seg_eager = {}, set()


class _CTrainGraph:
    """One minibatch shape on the C sequencer: the tensors the caller fills / reads and the sr_train_desc that binds
    them; plans, launch order and the CUDA graph live in libsr100 (sr_model_forward_backward, csrc/model.cu)."""

    def __init__(self, tr, NB, H, W):
        eng, dev = tr.engine, tr.engine.device
        self.NB, self.H, self.W = NB, H, W
        f32 = torch.float32
        self.x_in = torch.empty(NB, H, W, 3, device=dev, dtype=f32)
        self.y_true = torch.empty(NB, 4 * H, 4 * W, 3, device=dev, dtype=f32)
        self.out = torch.empty(NB, 4 * H, 4 * W, 3, device=dev, dtype=f32)
        self.loss_sum = torch.zeros(1, device=dev, dtype=torch.float64)
        self.n_local = NB * 16 * H * W * 3
        model = eng._ensure_model()
        need = eng.lib.sr_model_train_workspace_bytes(model, NB, H, W)
        self.workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        d = L.TrainDesc()
        d.NB, d.H, d.W = NB, H, W
        d.x, d.y, d.grads = self.x_in.data_ptr(), self.y_true.data_ptr(), tr.grads.data_ptr()
        d.loss_sum, d.pred = self.loss_sum.data_ptr(), self.out.data_ptr()
        d.workspace, d.workspace_bytes = self.workspace.data_ptr(), need
        self.desc = d
        info = L.ModelRunInfo()
        L.check(eng.lib.sr_model_train_info(model, C.byref(d), C.byref(info)))
        self.flops, self.launches = info.conv_flops, info.launches


class Trainer:
    """train_on_batch / evaluate for a kmodel.Model (Keras Model.train_on_batch semantics: returns the loss)."""

    def __init__(self, engine, lr=1e-4, beta_1=0.9, beta_2=0.999, epsilon=1e-7, exchange=None):
        if getattr(engine, "tf32", False):
            raise NotImplementedError("training runs on the bf16 engine (dgrad / wgrad kernels take bf16 operands); "
                                      "precision='tf32' is an inference mode")
        self.engine = engine
        self.lib = engine.lib
        self.lr, self.beta_1, self.beta_2, self.epsilon = float(lr), float(beta_1), float(beta_2), float(epsilon)
        dev = engine.device
        n = engine.n_params
        self.grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.t = 0
        self.workspace = torch.empty(self.lib.sr_wgrad_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.tail_d128 = torch.zeros(1, 1, NUMK, NUMK, dtype=torch.float32, device=dev)    # k=1 wgrad of the tail
        self.tail_colw = torch.zeros(1, 1, NUMK, NUMK, dtype=torch.float32, device=dev)    # B[j][ci] of its dgrad
        self.tail_colw_packed = torch.empty(self.lib.sr_packed_weight_bytes(1, NUMK), dtype=torch.uint8, device=dev)
        self.packed_t = {}      # name -> packed weights of the input-gradient conv
        self._pack_table_t = None
        self.comm_stream = None
        self._graphs = {}      # (NB, H, W) -> _TrainGraph, most recently used last (see graph())
        self.max_graphs = 3    # the full minibatch, the short last batch of a pass, the validation shape
        # "c": sr_model_forward_backward / sr_model_apply_gradients own the launch sequence; "python": the launch lists
        # of _TrainGraph (needed for the bucketed overlap and the per-kernel breakdown of tools/bench_train.py)
        self.sequencer = engine.sequencer
        self.exchange = None   # sr100.peer.Exchange: fused reduce-scatter + Adam + all-gather over NVLink peer memory
        self.sync_replicas()
        if self.sequencer != "c":
            self.repack_t()
        self._connect_exchange(exchange)

    def _connect_exchange(self, mode):
        """World > 1: the gradient exchange + optimizer step run as ONE kernel per rank over peer memory
        (csrc/exchange.cu) when every rank is on this node and the arenas can be mapped with CUDA IPC ('p2p', the
        default under NCCL); otherwise -- or with exchange='nccl' / SR100_EXCHANGE=nccl -- an all-reduce of the arena
        followed by the full-arena Adam launch."""
        import torch.distributed as tdist
        if not (tdist.is_available() and tdist.is_initialized()) or tdist.get_world_size() == 1:
            return
        mode = mode or os.environ.get("SR100_EXCHANGE") or ("p2p" if tdist.get_backend() == "nccl" else "nccl")
        if mode not in ("p2p", "nccl"):
            raise ValueError("exchange must be 'p2p' or 'nccl', got %r" % (mode,))
        if mode == "nccl":
            return
        from . import peer
        try:
            self.exchange = peer.connect(self.lib, self.grads, self.engine.param_arena)
        except RuntimeError as e:
            import warnings
            warnings.warn("sr100: %s; using the all-reduce exchange" % e, RuntimeWarning)

    def gather_optimizer_state(self):
        """With the peer-memory exchange every rank holds the Adam moments of its own shard only; this makes m and v
        whole on every rank again (before a checkpoint of the optimizer state or a replica re-sync)."""
        if self.exchange is None:
            return
        from .dist import all_reduce_sum_
        lo, hi = self.exchange.lo, self.exchange.hi
        for buf in (self.m, self.v):
            own = buf[lo:hi].clone()
            buf.zero_()
            buf[lo:hi] = own
            all_reduce_sum_(buf)

    # ------------------------------------------------------------------ views into the flat arenas
    def grad_w(self, name):
        ow, nw, _, _ = self.engine.param_slices[name]
        k, cin, cout = self.engine.ksize[name], (3 if name == "level1" else NUMK), self.engine.master[name][0].shape[3]
        return self.grads[ow:ow + nw].view(k, k, cin, cout)

    def grad_b(self, name):
        _, _, ob, nb = self.engine.param_slices[name]
        return self.grads[ob:ob + nb]

    def grads_dict(self):
        return {n: (self.grad_w(n).cpu().numpy(), self.grad_b(n).cpu().numpy()) for n, _, _, _ in self.engine.specs}

    def repack_t(self):
        """Weights of the input-gradient convs (180-degree rotated, cin <-> cout): one launch for all layers."""
        from .engine import PackTable
        self.engine.ensure_py_packed()
        if self._pack_table_t is None:
            for name, k, cin, cout in self.engine.specs:
                if cin == NUMK and name not in self.packed_t:
                    self.packed_t[name] = torch.empty(self.lib.sr_packed_weight_bytes(k, NUMK), dtype=torch.uint8,
                                                      device=self.engine.device)
            self._pack_table_t = PackTable(self.engine, self.packed_t, flip=True)
        self._pack_table_t.run()
        # tail dgrad as a 1x1 conv over the im2col'ed gradient: B[j = (ky,kx,co)][ci] = W[ky][kx][ci][co]
        tail = self.engine.specs[-1][0]
        w = self.engine.master[tail][0]                                   # [3,3,128,3]
        self.tail_colw.view(NUMK, NUMK)[:27].copy_(w.permute(0, 1, 3, 2).reshape(27, NUMK))
        L.check(self.lib.sr_pack_conv_weights(L.ptr(self.tail_colw), 1, NUMK, 0, L.ptr(self.tail_colw_packed),
                                              L.stream_ptr()))

    def sync_replicas(self):
        """Data parallelism needs identical replicas: broadcast rank 0's parameters and Adam state (m, v, t) to every
        rank (a no-op for one process).  Called at construction and after loading weights on rank 0 only; without it
        every rank would start from its own random initialisation and the averaged gradient would be applied to
        different weights (the ranks never meet again)."""
        import torch.distributed as tdist
        if not (tdist.is_available() and tdist.is_initialized()) or tdist.get_world_size() == 1:
            return False
        self.gather_optimizer_state()
        t = torch.tensor([float(self.t)], dtype=torch.float64, device=self.engine.device)
        for buf in (self.engine.param_arena, self.m, self.v, t):
            tdist.broadcast(buf, src=0)
        self.t = int(t.item())
        self.engine.repack()
        if self._pack_table_t is not None:
            self.repack_t()
        return True

    def graph(self, NB, H, W):
        """Buffers + launch lists for one minibatch shape.  A few shapes stay resident (LRU): `fit` alternates between
        the full batch, the short last batch of every pass over the data (img_utils._index_generator) and the
        validation shape, and rebuilding ~450 plans + a CUDA graph each time would stall every epoch.  If a new shape
        does not fit next to the cached ones, the others are dropped and it is tried once more."""
        key = (NB, H, W)
        g = self._graphs.pop(key, None)
        if g is None:
            while len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            make = _CTrainGraph if self.sequencer == "c" else _TrainGraph
            if make is _TrainGraph and self._pack_table_t is None:
                self.repack_t()
            try:
                g = make(self, NB, H, W)
            except torch.cuda.OutOfMemoryError:
                self._graphs.clear()
                torch.cuda.empty_cache()
                g = make(self, NB, H, W)
        self._graphs[key] = g      # most recently used last
        return g

    # ------------------------------------------------------------------ steps
    def _load(self, g, x, y):
        for dst, src in ((g.x_in, x), (g.y_true, y)):
            if isinstance(src, torch.Tensor):
                dst.copy_(src, non_blocking=True)
            else:
                dst.copy_(torch.from_numpy(np.ascontiguousarray(src, dtype=np.float32)), non_blocking=True)

    def forward_backward_device(self, g):
        """Forward + backward on the tensors already in g.x_in / g.y_true; gradients land in self.grads.  The ~450
        launches are a fixed sequence on fixed buffers: after one eager step they replay as one CUDA graph (at 32
        patches per GPU the step is otherwise launch-bound)."""
        if isinstance(g, _CTrainGraph):
            L.check(self.lib.sr_model_forward_backward(self.engine.model, C.byref(g.desc), L.stream_ptr()))
            return
        if self.engine.use_graphs and g.cuda_graph is not None:
            g.cuda_graph.replay()
            return

        def body(st):
            self.grads.zero_()
            g.loss_sum.zero_()
            for f in g.fwd:
                f(st)
            for f in g.bwd:
                f(st)

        body(L.stream_ptr())
        if self.engine.use_graphs and g.ran_eager and not torch.cuda.is_current_stream_capturing():
            try:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    body(L.stream_ptr())
                g.cuda_graph = gr
            except Exception as e:  # noqa: BLE001  (capture unsupported here: stay eager, but say so)
                import warnings
                warnings.warn("sr100: CUDA graph capture of the training step failed (%s: %s); running eagerly"
                              % (type(e).__name__, e), RuntimeWarning)
                self.engine.use_graphs = False
        g.ran_eager = True

    def apply_gradients(self, summed_over=None):
        """Data-parallel mean of the gradients (NCCL all-reduce of the flat arena) + fused Keras Adam + repack.
        summed_over: self.grads already holds the SUM over that many minibatch shards (a logical split run in one
        process: no collective, only the 1/shards scale) -- what the tests use to check the data-parallel algebra."""
        from .dist import all_reduce_sum_
        if self.exchange is not None and summed_over is None:
            self.t += 1
            ex = self.exchange
            args = (L.ptr(self.m), L.ptr(self.v), self.t, self.lr, self.beta_1, self.beta_2, self.epsilon, 1.0 / ex.world)
            if self.sequencer == "c":
                L.check(self.lib.sr_model_apply_gradients_exchange(self.engine.model, ex.handle, *args, L.stream_ptr()))
                if self.engine._py_packed:
                    self.engine.repack(c_model=False)
                    if self._pack_table_t is not None:
                        self.repack_t()
            else:
                L.check(self.lib.sr_exchange_adam_step(ex.handle, *args, 0, L.stream_ptr()))
                self.engine.repack()
                self.repack_t()
            return
        world = all_reduce_sum_(self.grads) if summed_over is None else int(summed_over)
        self.t += 1
        if self.sequencer == "c":      # Adam + every weight repack behind one entry point
            L.check(self.lib.sr_model_apply_gradients(self.engine.model, L.ptr(self.grads), L.ptr(self.m), L.ptr(self.v),
                                                      self.t, self.lr, self.beta_1, self.beta_2, self.epsilon,
                                                      1.0 / world, L.stream_ptr()))
            if self.engine._py_packed:           # something built Python-side plans too: keep their copies current
                self.engine.repack(c_model=False)
                if self._pack_table_t is not None:
                    self.repack_t()
            return
        L.check(self.lib.sr_adam_step(L.ptr(self.engine.param_arena), L.ptr(self.grads), L.ptr(self.m), L.ptr(self.v),
                                      self.engine.n_params, self.lr, self.beta_1, self.beta_2, self.epsilon, self.t,
                                      1.0 / world, L.stream_ptr()))
        self.engine.repack()
        self.repack_t()

    def _world(self):
        import torch.distributed as tdist
        return tdist.get_world_size() if (tdist.is_available() and tdist.is_initialized()) else 1

    def _run_segment(self, g, k, lo, hi):
        """Launches [lo, hi) of (zero, forward, backward) as segment k: eager the first time, then captured and
        replayed as its own CUDA graph (the all-reduce of a finished bucket is issued between segments)."""
        ops = g.all_ops
        gr = g.seg_graphs.get(k)
        if self.engine.use_graphs and gr is not None:
            gr.replay()
            return

        def body():
            st = L.stream_ptr()
            for f in ops[lo:hi]:
                f(st)

        body()
        if self.engine.use_graphs and k in g.seg_eager and not torch.cuda.is_current_stream_capturing():
            try:
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    body()
                g.seg_graphs[k] = gr
            except Exception as e:  # noqa: BLE001
                import warnings
                warnings.warn("sr100: CUDA graph capture of a training segment failed (%s: %s); running eagerly"
                              % (type(e).__name__, e), RuntimeWarning)
                self.engine.use_graphs = False
        g.seg_eager.add(k)

    def step_device(self, g, overlap=None):
        """One optimizer step on the tensors in g.x_in / g.y_true.  Default exchange: ONE all-reduce of the flat
        gradient arena after backward.  overlap=True (or SR100_OVERLAP_ALLREDUCE=1): the backward runs in three
        segments and the gradient bucket each segment completes (HR stage + tail; LR blocks 11-21; the rest) is
        all-reduced on a side stream while the next segment computes ("bucketed so the wgrad of early
        layers overlaps").  Measured on 2 B200s (profiles/r02_bench_n2_b.json): 89.4 ms overlapped vs 88.6 ms serial
        vs 87.8 ms without any exchange -- the NCCL kernel takes SMs away from the persistent one-CTA-per-SM conv /
        wgrad kernels it overlaps with, which costs more than the 0.8 ms it hides, so the serial exchange is the
        default and bench.py reports both.  World 1: no exchange at all."""
        import torch.distributed as tdist
        world = self._world()
        if overlap is None:
            overlap = os.environ.get("SR100_OVERLAP_ALLREDUCE", "0") == "1"
        if world == 1 or not overlap:
            self.forward_backward_device(g)
            self.apply_gradients()
            return
        if isinstance(g, _CTrainGraph):
            raise RuntimeError("the bucketed overlapped all-reduce needs the Python launch lists: "
                               "Engine(sequencer='python') or SR100_PY_SEQUENCE=1")
        if not hasattr(g, "all_ops"):
            zero = [lambda st: self.grads.zero_(), lambda st: g.loss_sum.zero_()]
            g.all_ops = zero + list(g.fwd) + list(g.bwd)
            base = len(zero) + len(g.fwd)
            g.segments, prev = [], 0
            for end, lo, hi in g.marks:
                g.segments.append((prev, base + end, lo, hi))
                prev = base + end
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(self.engine.device)
        main = torch.cuda.current_stream()
        works = []
        for k, (a, b, lo, hi) in enumerate(g.segments):
            self._run_segment(g, k, a, b)
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                works.append(tdist.all_reduce(self.grads[lo:hi], op=tdist.ReduceOp.SUM, async_op=True))
        for wk in works:
            wk.wait()                       # the main stream waits for the reductions
        main.wait_stream(self.comm_stream)
        self.apply_gradients(summed_over=world)

    def train_on_batch(self, x, y):
        x_shape = tuple(x.shape)
        if len(x_shape) != 4 or x_shape[-1] != 3 or tuple(y.shape) != (x_shape[0], 4 * x_shape[1], 4 * x_shape[2], 3):
            raise ValueError("Error when checking target: expected x (N,h,w,3) and y (N,4h,4w,3), got %s and %s"
                             % (x_shape, tuple(y.shape)))
        g = self.graph(*x_shape[:3])
        self._load(g, x, y)
        self.step_device(g)
        return self.last_loss(g)

    def last_loss(self, g):
        """MSE of the step just run, over the GLOBAL minibatch: sum of squared errors and element counts are summed
        over the ranks (what a single process would report for the concatenated batch)."""
        from .dist import all_reduce_sum_
        t = torch.stack([g.loss_sum[0], torch.tensor(float(g.n_local), dtype=torch.float64, device=g.loss_sum.device)])
        all_reduce_sum_(t)
        sse, n = t.tolist()
        return sse / n

    def comm_description(self):
        n = self.grads.numel() * 4
        if self.exchange is not None:
            return ("one fused kernel per rank over NVLink peer memory (CUDA IPC): reduce-scatter of the fp32 gradient "
                    "arena (%d bytes) in rank order + Keras-Adam on the rank's 1/%d shard + all-gather of the new "
                    "parameters into every rank's arena; no NCCL call on the data path" % (n, self.exchange.world))
        if os.environ.get("SR100_OVERLAP_ALLREDUCE", "0") == "1":
            return ("three all_reduce(sum) buckets of the fp32 gradient arena (%d bytes: HR stage + tail, LR blocks 11-21, "
                    "head + LR blocks 0-10), each issued on a side stream as soon as the backward segment that completes "
                    "it has been launched" % n)
        return "one all_reduce(sum) of the flat fp32 gradient arena (%d bytes) after backward" % n

    def evaluate(self, x, y):
        """(mse, categorical accuracy over the 3 colour channels) -- the compile(metrics=['accuracy']) pair."""
        g = self.graph(*tuple(x.shape)[:3])
        self._load(g, x, y)
        if isinstance(g, _CTrainGraph):
            g.out.copy_(self.engine.forward_device(g.x_in))
        else:
            st = L.stream_ptr()
            for f in g.fwd:
                f(st)
        d = g.out - g.y_true
        acc = (g.out.argmax(dim=-1) == g.y_true.argmax(dim=-1)).float().mean()
        return float((d * d).mean().item()), float(acc.item())

    def graph_ready(self, g):
        """True when the next forward_backward_device(g) replays a captured CUDA graph."""
        if isinstance(g, _CTrainGraph):
            info = L.ModelRunInfo()
            L.check(self.lib.sr_model_train_info(self.engine.model, C.byref(g.desc), C.byref(info)))
            return bool(info.graph_replay)
        return g.cuda_graph is not None

    def step_flops(self, g):
        return g.flops if isinstance(g, _CTrainGraph) else g.fwd_flops + g.bwd_flops
