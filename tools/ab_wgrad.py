"""A/B of two builds of libsr100.so on the filter-gradient kernel, in ONE process (dev tool).

    python tools/ab_wgrad.py --old image-enhance-keras_b200/lib_ab/libsr100_old.so [--out gpurun_out/ab_wgrad.json]

Both libraries run the same plans on the same tensors.  A change that only moves the MMA issue loop between register
files issues the same tcgen05.mma sequence, so the gradients must be BIT-identical (checked on ragged and multi-segment
shapes, with scale / accumulate), and each result is also compared with torch's conv2d_weight.  Timing: the perf shapes
of tools/probe_wgrad.py, the two libraries interleaved, CUDA events.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))

CHECK = [dict(NB=2, H=12, W=48, k=1), dict(NB=2, H=12, W=48, k=3), dict(NB=2, H=12, W=48, k=5),
         dict(NB=1, H=3, W=16, k=3), dict(NB=3, H=33, W=50, k=5), dict(NB=2, H=20, W=96, k=3, scale=0.1, accumulate=1),
         dict(NB=2, H=24, W=192, k=5), dict(NB=1, H=9, W=384, k=3), dict(NB=40, H=48, W=48, k=5),
         dict(NB=7, H=10, W=32, k=3), dict(NB=5, H=9, W=16, k=5), dict(NB=3, H=17, W=48, k=3),
         dict(NB=200, H=5, W=20, k=5), dict(NB=2, H=192, W=192, k=5)]
PERF = [dict(NB=256, H=48, W=48, k=3, iters=10), dict(NB=256, H=48, W=48, k=5, iters=10),
        dict(NB=32, H=192, W=192, k=3, iters=5), dict(NB=32, H=192, W=192, k=5, iters=5),
        dict(NB=256, H=192, W=192, k=3, iters=4), dict(NB=256, H=192, W=192, k=5, iters=4)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--old", required=True)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ab_wgrad.json"))
    a = ap.parse_args()
    t_start = time.time()
    import torch
    from sr100 import _lib as L
    new = L.require_device()
    old = C.CDLL(os.path.abspath(a.old))
    for name in ("sr_wgrad_workspace_bytes", "sr_wgrad_plan_create", "sr_wgrad_plan_run", "sr_wgrad_plan_destroy",
                 "sr_wgrad_plan_info"):
        getattr(old, name).restype, getattr(old, name).argtypes = L.SIGNATURES[name]
    res = dict(import_s=round(time.time() - t_start, 1), check=[], perf=[])
    dev = "cuda"

    def plan(lib, x, g, dw, ws, cs):
        d = L.WgradDesc()
        d.x_bf16, d.g_bf16 = x.data_ptr(), g.data_ptr()
        d.NB, d.H, d.W, d.ksize = cs["NB"], cs["H"], cs["W"], cs["k"]
        d.scale, d.accumulate = cs.get("scale", 1.0), cs.get("accumulate", 0)
        d.dw_hwio, d.workspace, d.workspace_bytes = dw.data_ptr(), ws.data_ptr(), ws.numel()
        p = C.c_void_p()
        rc = lib.sr_wgrad_plan_create(C.byref(d), C.byref(p))
        assert rc == 0, rc
        return p

    ws = torch.empty(new.sr_wgrad_workspace_bytes(), dtype=torch.uint8, device=dev)
    ok_all = True
    for i, cs in enumerate(CHECK):
        torch.manual_seed(i)
        NB, H, W, k = cs["NB"], cs["H"], cs["W"], cs["k"]
        x = (torch.randn(NB, H, W, 128, device=dev) * 0.5).to(torch.bfloat16).contiguous()
        g = (torch.randn(NB, H, W, 128, device=dev) * 0.5).to(torch.bfloat16).contiguous()
        dw0 = torch.randn(k, k, 128, 128, device=dev)
        outs = []
        for lib in (new, old):
            dw = dw0.clone()
            p = plan(lib, x, g, dw, ws, cs)
            assert lib.sr_wgrad_plan_run(p, L.stream_ptr()) == 0
            torch.cuda.synchronize()
            lib.sr_wgrad_plan_destroy(p)
            outs.append(dw)
        want = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2).contiguous(), (128, 128, k, k),
                                           g.float().permute(0, 3, 1, 2).contiguous(), padding=k // 2)
        want = want.permute(2, 3, 1, 0).contiguous() * cs.get("scale", 1.0)
        if cs.get("accumulate"):
            want = want + dw0
        err = float((outs[0] - want).abs().max())
        ref = float(want.abs().max())
        same = bool(torch.equal(outs[0], outs[1]))
        ok = same and err <= 2e-3 * max(ref, 1.0)
        ok_all = ok_all and ok
        res["check"].append(dict(cs, bit_identical_to_old=same, max_err_vs_torch=err, max_ref=ref, ok=ok))
    res["all_checks_ok"] = ok_all
    print(json.dumps(dict(all_checks_ok=ok_all, n=len(CHECK), t=round(time.time() - t_start, 1))), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)

    for cs in PERF:
        NB, H, W, k = cs["NB"], cs["H"], cs["W"], cs["k"]
        x = (torch.randn(NB, H, W, 128, device=dev) * 0.5).to(torch.bfloat16).contiguous()
        g = (torch.randn(NB, H, W, 128, device=dev) * 0.5).to(torch.bfloat16).contiguous()
        dw = torch.zeros(k, k, 128, 128, device=dev)
        plans = {"new": (new, plan(new, x, g, dw, ws, cs)), "old": (old, plan(old, x, g, dw, ws, cs))}
        flops = 2.0 * NB * H * W * k * k * 128 * 128
        rec = dict(cs)
        ms = {"new": [], "old": []}
        for rep in range(3):
            for name in ("old", "new"):
                lib, p = plans[name]
                lib.sr_wgrad_plan_run(p, L.stream_ptr())
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(cs["iters"]):
                    lib.sr_wgrad_plan_run(p, L.stream_ptr())
                e1.record()
                torch.cuda.synchronize()
                ms[name].append(e0.elapsed_time(e1) / cs["iters"])
        for name in ("old", "new"):
            best = min(ms[name])
            rec[name + "_ms"] = round(best, 4)
            rec[name + "_tflops"] = round(flops / best / 1e9, 1)
            plans[name][0].sr_wgrad_plan_destroy(plans[name][1])
        rec["speedup"] = round(rec["old_ms"] / rec["new_ms"], 4)
        res["perf"].append(rec)
        print(json.dumps(rec), flush=True)
        del x, g
    res["total_s"] = round(time.time() - t_start, 1)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    sys.exit(0 if ok_all else 4)


if __name__ == "__main__":
    main()
