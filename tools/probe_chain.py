"""sr_conv_chain (a sequence of convolutions in one persistent launch, grid barrier between phases) against the same
convolutions as single plans -- eager and replayed from a CUDA graph -- and against the chain with every convolution
in phase 0 (no barriers: wrong results, shows what the barriers cost).  Shapes of BASELINE config 1's LR stage.

    python tools/probe_chain.py                       # timings, one line per case
    SR100_LIB=image-enhance-keras_b200/lib_dev/libsr100.so python tools/probe_chain.py --timeline
        (development build, make DEV=1: per-phase globaltimer stamps of a light-block chain, min / median / max over CTAs)
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))
from sr100 import _lib as L  # noqa: E402

lib = L.require_device()
dev, bf = "cuda", torch.bfloat16
rng = np.random.default_rng(0)


def packed(k):
    w = torch.from_numpy((rng.normal(0, 0.02, size=(k, k, 128, 128))).astype(np.float32)).to(dev)
    pk = torch.empty(lib.sr_packed_weight_bytes(k, 128), dtype=torch.uint8, device=dev)
    L.check(lib.sr_pack_conv_weights(L.ptr(w), k, 128, 0, L.ptr(pk), L.stream_ptr()))
    return pk


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def build(NB, H, W, n53, nlight):
    """Descriptors + phases of n53 5/3 blocks followed by nlight light blocks (the engine's LR stage is 16 + 6)."""
    x0 = torch.from_numpy(rng.normal(0, 1, size=(NB, H, W, 128)).astype(np.float32)).to(dev)
    bias = torch.from_numpy(rng.normal(0, 0.05, size=128).astype(np.float32)).to(dev)
    s32 = x0.clone()
    s = s32.to(bf)
    t1, t2 = torch.zeros_like(s), torch.zeros_like(s)
    keep = [x0, bias, s32, s, t1, t2]
    descs, phases, ph = [], [], 0

    def mk(srcs, out16, relu=0, alpha=1.0, beta=0.0, res=False):
        d = L.ConvDesc()
        d.nsrc, d.NB, d.H, d.W, d.cin, d.cout = len(srcs), NB, H, W, 128, 128
        for i, (t, wk, k) in enumerate(srcs):
            d.in_[i], d.wpacked[i], d.ksize[i] = t.data_ptr(), wk.data_ptr(), k
            keep.append(wk)
        d.bias, d.alpha, d.beta, d.relu = bias.data_ptr(), alpha, beta, relu
        d.out_bf16 = out16.data_ptr()
        if res:
            d.res_f32, d.out_f32 = s32.data_ptr(), s32.data_ptr()
        d.nacc, d.pair = 2, 1
        return d

    for _ in range(n53):
        a, c, b, d_ = packed(3), packed(5), packed(5), packed(3)
        descs += [mk([(s, a, 3)], t1, relu=1), mk([(s, c, 5)], t2, relu=1)]
        phases += [ph, ph]
        ph += 1
        descs += [mk([(t1, b, 5), (t2, d_, 3)], s, alpha=0.1, beta=0.9, res=True)]
        phases += [ph]
        ph += 1
    for _ in range(nlight):
        a, b = packed(3), packed(3)
        descs += [mk([(s, a, 3)], t1, relu=1)]
        phases += [ph]
        ph += 1
        descs += [mk([(t1, b, 3)], s, alpha=0.1, beta=1.0, res=True)]
        phases += [ph]
        ph += 1
    return descs, phases, ph, keep


def chain_of(descs, phases):
    arr = (L.ConvDesc * len(descs))(*descs)
    pa = (C.c_int * len(phases))(*phases)
    h = C.c_void_p()
    L.check(lib.sr_conv_chain_create(arr, pa, len(descs), C.byref(h)))
    return h


def timings(NB, H, W, n53, nlight):
    descs, phases, ph, keep = build(NB, H, W, n53, nlight)
    out = {}
    for name, pl in (("chain", phases), ("chain_no_barriers", [0] * len(phases))):
        h = chain_of(descs, pl)
        out[name] = timeit(lambda: L.check(lib.sr_conv_chain_run(h, L.stream_ptr())))
        lib.sr_conv_chain_destroy(h)
    plans = []
    for d in descs:
        p = C.c_void_p()
        L.check(lib.sr_conv_plan_create(C.byref(d), C.byref(p)))
        plans.append(p)

    def run_plans():
        for p in plans:
            L.check(lib.sr_conv_plan_run(p, L.stream_ptr()))

    out["plans_eager"] = timeit(run_plans)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run_plans()
    out["plans_graph"] = timeit(g.replay)
    for p in plans:
        lib.sr_conv_plan_destroy(p)
    print("shape %s  %d 5/3 blocks + %d light blocks = %d phases, us: %s;  per phase: chain %.1f, plans from a graph %.1f"
          % ((NB, H, W), n53, nlight, ph, {k: round(v, 1) for k, v in out.items()}, out["chain"] / ph,
             out["plans_graph"] / ph), flush=True)


def timeline():
    if not lib.sr_dev_switches():
        raise SystemExit("--timeline needs the development build (make -C image-enhance-keras_b200/csrc DEV=1, SR100_LIB=...)")
    descs, phases, ph, keep = build(1, 128, 128, 0, 4)
    h = chain_of(descs, phases)
    info = L.ConvPlanInfo()
    lib.sr_conv_chain_info(h, C.byref(info))
    G = info.grid
    for _ in range(3):
        L.check(lib.sr_conv_chain_run(h, L.stream_ptr()))
    tl = torch.zeros(G * ph * 8, dtype=torch.int64, device=dev)
    L.check(lib.sr_dev_set_timeline(C.c_void_p(tl.data_ptr())))
    L.check(lib.sr_conv_chain_run(h, L.stream_ptr()))
    torch.cuda.synchronize()
    L.check(lib.sr_dev_set_timeline(None))
    t = tl.cpu().numpy().reshape(G, ph, 8).astype(np.float64)
    t0 = t[t > 0].min()
    t = np.where(t > 0, (t - t0) / 1e3, np.nan)
    names = ["barrier passed", "first strip ready", "last strip ready", "mma committed", "epilogue sees acc",
             "epilogue done", "fenced, about to arrive"]
    print("light-block chain on one 128x128 image: grid %d, %d phases, %d weight stages; microseconds since the first stamp"
          % (G, ph, info.num_wstages))
    for p_ in range(ph):
        row = []
        for i, nm in enumerate(names):
            col = t[:, p_, i]
            if np.all(np.isnan(col)):
                continue
            row.append("%s %.1f / %.1f / %.1f" % (nm, np.nanmin(col), np.nanmedian(col), np.nanmax(col)))
        print("phase %d (%s): %s" % (p_, "k3 + relu" if p_ % 2 == 0 else "k3, fp32 residual", "; ".join(row)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--timeline", action="store_true")
    a = ap.parse_args()
    if a.timeline:
        timeline()
    else:
        for cfg in [(1, 128, 128, 0, 4), (1, 128, 128, 4, 0), (1, 128, 128, 16, 6), (1, 64, 64, 16, 6)]:
            timings(*cfg)
