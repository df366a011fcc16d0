"""-m gpu: randomised geometry sweep of the tensor-core conv (segmenting, halo strips, ragged widths, odd batches
under the CTA-pair kernel, compute extents, every epilogue specialisation) against the CUDA-core direct conv, which
tests/test_gpu_conv.py pins to the CPU oracle.  Seeds are fixed: the sweep is deterministic."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(rng):
    k = [int(rng.choice([1, 3, 5]))]
    if rng.random() < 0.35:
        k.append(int(rng.choice([3, 5])))
    NB = int(rng.integers(1, 6))
    H = int(rng.integers(1, 34))
    W = int(rng.choice([int(rng.integers(1, 40)), int(rng.integers(40, 140)), int(rng.integers(140, 300))]))
    mode = str(rng.choice(["plain", "res32", "res16", "mask", "generic"]))
    comp = None
    if rng.random() < 0.4 and H > 2 and W > 2:
        comp = (int(rng.integers(1, H + 1)), int(rng.integers(1, W + 1)))
    return dict(k=k, NB=NB, H=H, W=W, mode=mode, comp=comp, pair=int(rng.random() < 0.7), relu=int(rng.random() < 0.5))


@pytest.mark.parametrize("seed", range(6))
def test_conv_geometry_sweep(lib, seed):
    from sr100 import _lib as L
    rng = np.random.default_rng(1000 + seed)
    torch.manual_seed(seed)
    for _ in range(8):
        cs = _case(rng)
        NB, H, W = cs["NB"], cs["H"], cs["W"]
        d = L.ConvDesc()
        d.nsrc = len(cs["k"])
        keep, acc = [], torch.zeros(NB, H, W, 128, device="cuda")
        for s, k in enumerate(cs["k"]):
            x = (torch.randn(NB, H, W, 128, device="cuda") * 0.5).to(torch.bfloat16)
            w = torch.randn(k, k, 128, 128, device="cuda") / (k * k * 128) ** 0.5
            pk = torch.empty(lib.sr_packed_weight_bytes(k, 128), dtype=torch.uint8, device="cuda")
            L.check(lib.sr_pack_conv_weights(L.ptr(w), k, 128, 0, L.ptr(pk), L.stream_ptr()))
            d.in_[s], d.wpacked[s], d.ksize[s] = x.data_ptr(), pk.data_ptr(), k
            tmp = torch.empty(NB, H, W, 128, device="cuda")
            L.check(lib.sr_conv2d_direct(L.ptr(x), 1, L.ptr(w), 1, None, NB, H, W, 128, 128, k, 1, 0, 0, 0, L.ptr(tmp),
                                         L.stream_ptr()))
            acc += tmp
            keep += [x, w, pk]
        bias = torch.randn(128, device="cuda") * 0.1
        alpha, beta = 0.5, 0.0
        want = acc
        res = mask = None
        d.NB, d.H, d.W, d.cin, d.cout = NB, H, W, 128, 128
        d.bias = bias.data_ptr()
        out_f = torch.full((NB, H, W, 128), float("nan"), device="cuda")
        out_b = torch.zeros(NB, H, W, 128, device="cuda", dtype=torch.bfloat16)
        d.out_bf16 = out_b.data_ptr()
        if cs["mode"] in ("res32", "generic"):
            res = torch.randn(NB, H, W, 128, device="cuda")
            d.res_f32, beta = res.data_ptr(), 0.9
            d.out_f32 = out_f.data_ptr()
        elif cs["mode"] == "res16":
            res = torch.randn(NB, H, W, 128, device="cuda").to(torch.bfloat16)
            d.res_bf16, beta = res.data_ptr(), 0.9
        if cs["mode"] in ("mask", "generic"):
            mask = torch.randn(NB, H, W, 128, device="cuda").to(torch.bfloat16)
            d.relu_mask_bf16 = mask.data_ptr()
        d.alpha, d.beta, d.relu = alpha, beta, cs["relu"]
        d.a_mode, d.nacc, d.pair = 0, 2, cs["pair"]
        if cs["comp"]:
            d.comp_h, d.comp_w = cs["comp"]
        plan = C.c_void_p()
        L.check(lib.sr_conv_plan_create(C.byref(d), C.byref(plan)))
        L.check(lib.sr_conv_plan_run(plan, L.stream_ptr()))
        torch.cuda.synchronize()
        lib.sr_conv_plan_destroy(plan)
        # reference with the kernel's rounding points: accumulator -> bf16, then fp32 alpha / bias / residual / act
        ref = alpha * want.to(torch.bfloat16).float() + alpha * bias
        if res is not None:
            ref = ref + beta * res.float()
        if cs["relu"]:
            ref = ref.clamp_min(0)
        if mask is not None:
            ref = torch.where(mask.float() > 0, ref, torch.zeros_like(ref))
        ch, cw = cs["comp"] if cs["comp"] else (H, W)
        got = out_b.float()[:, :ch, :cw]
        r = ref[:, :ch, :cw]
        tol = 2.0 ** -7 * max(1.0, float(r.abs().max()))          # bf16 output rounding (+ one accumulator ulp)
        assert float((got - r).abs().max()) <= tol, cs
        if d.out_f32:
            assert float((out_f[:, :ch, :cw] - r).abs().max()) <= tol, cs
        if cs["comp"]:
            assert (out_b[:, ch:] == 0).all() and (out_b[:, :, cw:] == 0).all(), cs
