"""Sweep of the filter-gradient kernel's CTA split between tap groups (dev tool; needs the development build).

    SR100_LIB=image-enhance-keras_b200/lib_dev/libsr100.so python tools/sweep_wgrad_split.py [--out gpurun_out/x.json]

Every tap group streams the same X and G rows, so the part of a group's time that is operand ingest does not depend on
its tap count; SR100_WGRAD_COSTFLOOR puts a floor under the group costs the CTAs are shared out by (0 = by MMA work,
>= 3 = equal CTAs per group for the 5x5).  The environment is read when a plan is created, so all variants run in one
process on the same tensors, interleaved; the result of every variant is checked against torch on a small shape.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))

VARIANTS = [dict(floor="0"), dict(floor="2.75"), dict(floor="3.0"), dict(floor="3.0", pcost="1.75")]
PERF = [dict(NB=256, H=48, W=48, k=5, iters=10), dict(NB=32, H=192, W=192, k=5, iters=5),
        dict(NB=256, H=192, W=192, k=5, iters=3)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_wgrad_split.json"))
    a = ap.parse_args()
    t0 = time.time()
    import torch
    from sr100 import _lib as L
    lib = L.require_device()
    assert lib.sr_dev_switches() == 1, "needs the development build (make DEV=1, SR100_LIB=.../lib_dev/libsr100.so)"
    dev = "cuda"
    ws = torch.empty(lib.sr_wgrad_workspace_bytes(), dtype=torch.uint8, device=dev)

    def plan(x, g, dw, cs, v):
        os.environ["SR100_WGRAD_COSTFLOOR"] = v["floor"]
        if "pcost" in v:
            os.environ["SR100_WGRAD_PCOST"] = v["pcost"]
        else:
            os.environ.pop("SR100_WGRAD_PCOST", None)
        d = L.WgradDesc()
        d.x_bf16, d.g_bf16 = x.data_ptr(), g.data_ptr()
        d.NB, d.H, d.W, d.ksize = cs["NB"], cs["H"], cs["W"], cs["k"]
        d.scale, d.accumulate = 1.0, 0
        d.dw_hwio, d.workspace, d.workspace_bytes = dw.data_ptr(), ws.data_ptr(), ws.numel()
        p = C.c_void_p()
        L.check(lib.sr_wgrad_plan_create(C.byref(d), C.byref(p)))
        return p

    res = dict(check=[], perf=[])
    cs = dict(NB=3, H=33, W=50, k=5)
    torch.manual_seed(0)
    x = (torch.randn(3, 33, 50, 128, device=dev) * 0.5).to(torch.bfloat16)
    g = (torch.randn(3, 33, 50, 128, device=dev) * 0.5).to(torch.bfloat16)
    want = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2).contiguous(), (128, 128, 5, 5),
                                       g.float().permute(0, 3, 1, 2).contiguous(), padding=2).permute(2, 3, 1, 0)
    for v in VARIANTS:
        dw = torch.zeros(5, 5, 128, 128, device=dev)
        p = plan(x, g, dw, cs, v)
        L.check(lib.sr_wgrad_plan_run(p, L.stream_ptr()))
        torch.cuda.synchronize()
        lib.sr_wgrad_plan_destroy(p)
        err = float((dw - want).abs().max())
        res["check"].append(dict(v, max_err=err, ok=err <= 2e-3 * max(float(want.abs().max()), 1.0)))
    print(json.dumps(res["check"]), flush=True)

    for cs in PERF:
        NB, H, W, k = cs["NB"], cs["H"], cs["W"], cs["k"]
        x = (torch.randn(NB, H, W, 128, device=dev) * 0.5).to(torch.bfloat16)
        g = (torch.randn(NB, H, W, 128, device=dev) * 0.5).to(torch.bfloat16)
        dw = torch.zeros(k, k, 128, 128, device=dev)
        plans = [plan(x, g, dw, cs, v) for v in VARIANTS]
        flops = 2.0 * NB * H * W * k * k * 128 * 128
        ms = [[] for _ in VARIANTS]
        for rep in range(3):
            for i, p in enumerate(plans):
                lib.sr_wgrad_plan_run(p, L.stream_ptr())
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(cs["iters"]):
                    lib.sr_wgrad_plan_run(p, L.stream_ptr())
                e1.record()
                torch.cuda.synchronize()
                ms[i].append(e0.elapsed_time(e1) / cs["iters"])
        rec = dict(cs, variants=[dict(v, ms=round(min(m), 4), tflops=round(flops / min(m) / 1e9, 1))
                                 for v, m in zip(VARIANTS, ms)])
        for p in plans:
            lib.sr_wgrad_plan_destroy(p)
        res["perf"].append(rec)
        print(json.dumps(rec), flush=True)
        del x, g
    res["total_s"] = round(time.time() - t0, 1)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
