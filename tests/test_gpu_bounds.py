"""-m gpu: no kernel writes outside its output tensors (compute-sanitizer is not available on the pool, so the outputs
are carved out of larger buffers whose guard bands must stay untouched) -- odd batch sizes under the CTA-pair kernel,
ragged widths, cropped compute extents, the sub-pixel scatter, the wgrad drain, bilinear crop and the stitch."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096


def _guarded(shape, dtype, fill):
    n = int(np.prod(shape))
    buf = torch.full((n + 2 * GUARD,), fill, device="cuda", dtype=dtype)
    return buf, buf[GUARD:GUARD + n].view(*shape)


def _intact(buf, n, fill):
    g0, g1 = buf[:GUARD], buf[GUARD + n:]
    if isinstance(fill, float) and fill != fill:
        return bool(torch.isnan(g0).all() and torch.isnan(g1).all())
    return bool((g0 == fill).all() and (g1 == fill).all())


@pytest.mark.parametrize("NB,H,W,ks,comp", [(3, 9, 50, (5,), None), (1, 7, 96, (3,), None), (5, 20, 33, (5, 3), None),
                                            (3, 30, 40, (3,), (17, 23)), (2, 12, 200, (5,), (12, 150))])
def test_conv_outputs_stay_in_bounds(lib, NB, H, W, ks, comp):
    from sr100 import _lib as L
    torch.manual_seed(NB * 100 + W)
    n = NB * H * W * 128
    ob_buf, ob = _guarded((NB, H, W, 128), torch.bfloat16, 7.0)
    of_buf, of = _guarded((NB, H, W, 128), torch.float32, float("nan"))
    keep = []
    d = L.ConvDesc()
    d.nsrc = len(ks)
    for s, k in enumerate(ks):
        x = (torch.randn(NB, H, W, 128, device="cuda") * 0.5).to(torch.bfloat16)
        w = torch.randn(k, k, 128, 128, device="cuda") / (k * k * 128) ** 0.5
        pk = torch.empty(lib.sr_packed_weight_bytes(k, 128), dtype=torch.uint8, device="cuda")
        L.check(lib.sr_pack_conv_weights(L.ptr(w), k, 128, 0, L.ptr(pk), L.stream_ptr()))
        d.in_[s], d.wpacked[s], d.ksize[s] = x.data_ptr(), pk.data_ptr(), k
        keep += [x, w, pk]
    res = torch.randn(NB, H, W, 128, device="cuda")
    d.NB, d.H, d.W, d.cin, d.cout = NB, H, W, 128, 128
    d.alpha, d.beta, d.relu = 0.1, 0.9, 0
    d.res_f32 = res.data_ptr()
    d.out_bf16, d.out_f32 = ob.data_ptr(), of.data_ptr()
    d.a_mode, d.nacc, d.pair = 0, 2, 1
    if comp:
        d.comp_h, d.comp_w = comp
    plan = C.c_void_p()
    L.check(lib.sr_conv_plan_create(C.byref(d), C.byref(plan)))
    L.check(lib.sr_conv_plan_run(plan, L.stream_ptr()))
    torch.cuda.synchronize()
    lib.sr_conv_plan_destroy(plan)
    assert _intact(ob_buf, n, 7.0) and _intact(of_buf, n, float("nan"))
    ch, cw = comp if comp else (H, W)
    assert not torch.isnan(of[:, :ch, :cw]).any()                       # the compute region is fully written
    if comp:                                                            # ... and nothing outside it
        assert torch.isnan(of[:, ch:]).all() and torch.isnan(of[:, :, cw:]).all()
        assert (ob[:, ch:] == 7.0).all() and (ob[:, :, cw:] == 7.0).all()


def test_wgrad_bilinear_stitch_stay_in_bounds(lib):
    from sr100 import _lib as L
    from sr100 import ops
    torch.manual_seed(1)
    # wgrad: dw and the workspace
    x = torch.randn(3, 11, 48, 128, device="cuda").to(torch.bfloat16)
    g = torch.randn(3, 11, 48, 128, device="cuda").to(torch.bfloat16)
    dw_buf, dw = _guarded((5, 5, 128, 128), torch.float32, float("nan"))
    nws = lib.sr_wgrad_workspace_bytes()
    ws_buf = torch.full((nws + 2 * GUARD,), 9, device="cuda", dtype=torch.uint8)
    d = L.WgradDesc()
    d.x_bf16, d.g_bf16 = x.data_ptr(), g.data_ptr()
    d.NB, d.H, d.W, d.ksize = 3, 11, 48, 5
    d.scale, d.accumulate = 1.0, 0
    d.dw_hwio, d.workspace, d.workspace_bytes = dw.data_ptr(), ws_buf[GUARD:].data_ptr(), nws
    plan = C.c_void_p()
    L.check(lib.sr_wgrad_plan_create(C.byref(d), C.byref(plan)))
    L.check(lib.sr_wgrad_plan_run(plan, L.stream_ptr()))
    torch.cuda.synchronize()
    lib.sr_wgrad_plan_destroy(plan)
    assert _intact(dw_buf, dw.numel(), float("nan")) and not torch.isnan(dw).any()
    assert (ws_buf[:GUARD] == 9).all() and (ws_buf[GUARD + nws:] == 9).all()
    # bilinear crop with a source gather
    src = torch.randn(4, 10, 12, 128, device="cuda")
    idx = torch.tensor([3, 1], dtype=torch.int32, device="cuda")
    o_buf, o = _guarded((2, 24, 32, 128), torch.bfloat16, 7.0)
    L.check(lib.sr_bilinear4_crop_fwd(L.ptr(src), 0, L.ptr(idx), 2, 10, 12, 128, 24, 32, L.ptr(o), None, L.stream_ptr()))
    torch.cuda.synchronize()
    assert _intact(o_buf, o.numel(), 7.0)
    full = ops.bilinear4(src[[3, 1]], out_dtype=torch.bfloat16)
    assert torch.equal(o, full[:, :24, :32].contiguous())
    # stitch into a cropped image whose width is not a multiple of 4 pixels
    patches = torch.rand(6, 384, 384, 3, device="cuda")
    u_buf, u = _guarded((4 * 70, 4 * 101, 3), torch.uint8, 5)
    L.check(lib.sr_patch_stitch(L.ptr(patches), 2, 3, 96, 96, 64, 4, 70, 101, 255.0, None, L.ptr(u), L.stream_ptr()))
    torch.cuda.synchronize()
    assert _intact(u_buf, u.numel(), 5)
