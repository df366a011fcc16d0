"""Helpers shared by the -m gpu tests (which call the product through the C ABI and check it against the
CPU oracle)."""
import ctypes as C

import numpy as np
import torch


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def run_tc_conv(lib, xs, ws, bias, relu=0, alpha=1.0, beta=0.0, res=None, res_dtype="f32", cout=128, a_mode=0,
                nacc=4, pair=0):
    """xs: list of float32 NHWC arrays (bf16-representable), ws: list of HWIO float32.  Returns (out_f32, out_bf16)."""
    from sr100 import _lib as L
    dev = "cuda"
    NB, H, W, _ = xs[0].shape
    keep = []
    d = L.ConvDesc()
    d.nsrc = len(xs)
    for s, (x, w) in enumerate(zip(xs, ws)):
        k = w.shape[0]
        xd = torch.from_numpy(x).to(dev).to(torch.bfloat16).contiguous()
        wd = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32)).to(dev)
        pk = torch.empty(lib.sr_packed_weight_bytes(k, cout), dtype=torch.uint8, device=dev)
        L.check(lib.sr_pack_conv_weights(L.ptr(wd), k, cout, 0, L.ptr(pk), L.stream_ptr()))
        d.in_[s], d.wpacked[s], d.ksize[s] = xd.data_ptr(), pk.data_ptr(), k
        keep += [xd, wd, pk]
    d.NB, d.H, d.W, d.cin, d.cout = NB, H, W, 128, cout
    bd = torch.from_numpy(np.ascontiguousarray(bias, dtype=np.float32)).to(dev) if bias is not None else None
    d.bias = bd.data_ptr() if bd is not None else None
    d.alpha, d.beta, d.relu = alpha, beta, relu
    rd = None
    if res is not None:
        rd = torch.from_numpy(res).to(dev)
        if res_dtype == "bf16":
            rd = rd.to(torch.bfloat16)
            d.res_bf16 = rd.data_ptr()
        else:
            d.res_f32 = rd.data_ptr()
    of = torch.full((NB, H, W, cout), float("nan"), device=dev)
    ob = torch.zeros(NB, H, W, cout, device=dev, dtype=torch.bfloat16)
    d.out_f32, d.out_bf16 = of.data_ptr(), ob.data_ptr()
    d.a_mode, d.nacc, d.pair = a_mode, nacc, pair
    plan = C.c_void_p()
    L.check(lib.sr_conv_plan_create(C.byref(d), C.byref(plan)))
    L.check(lib.sr_conv_plan_run(plan, L.stream_ptr()))
    torch.cuda.synchronize()
    lib.sr_conv_plan_destroy(plan)
    return of.cpu().numpy(), ob.float().cpu().numpy()


def oracle_conv(xs, ws, bias, relu=0, alpha=1.0, beta=0.0, res=None):
    """CPU restatement (Keras Conv2D semantics via the oracle's torch conv) with bf16-rounded operands."""
    import torch.nn.functional as F
    acc = None
    for x, w in zip(xs, ws):
        k = w.shape[0]
        y = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(bf16_round(w)).permute(3, 2, 0, 1),
                     padding=(k - 1) // 2)
        acc = y if acc is None else acc + y
    if bias is not None:
        acc = acc + torch.from_numpy(np.asarray(bias, dtype=np.float32)).view(1, -1, 1, 1)
    out = alpha * acc.permute(0, 2, 3, 1)
    if res is not None:
        out = out + beta * torch.from_numpy(res)
    if relu:
        out = out.clamp_min(0)
    return out.numpy()
