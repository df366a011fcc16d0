"""GPU probe for the tcgen05 wgrad kernel: each case in its own subprocess under a timeout, checked against
cuDNN's filter gradient (dev tool only; the tests use the CPU oracle), TFLOP/s for the big cases.

    python tools/probe_wgrad.py [--only substr] [--out gpurun_out/probe_wgrad.jsonl]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))

CASES = []


def case(**kw):
    d = dict(NB=2, H=12, W=48, k=3, iters=0, check=True, scale=1.0, accumulate=0)
    d.update(kw)
    CASES.append(d)


case(name="k1_w48", k=1)
case(name="k3_w48", k=3)
case(name="k5_w48", k=5)
case(name="k3_w16_h3", k=3, W=16, H=3, NB=1)
case(name="k5_w50_h33", k=5, W=50, H=33, NB=3)
case(name="k3_w96_scale_acc", k=3, W=96, H=20, scale=0.1, accumulate=1)
case(name="k5_w192_h24", k=5, W=192, H=24, NB=2)
case(name="k3_w384_h9", k=3, W=384, H=9, NB=1)
case(name="k5_many_units", k=5, W=48, H=48, NB=40)
case(name="k3_w32_nb7", k=3, W=32, H=10, NB=7)
case(name="k5_w16_nb5", k=5, W=16, H=9, NB=5)
case(name="k3_w48_nb3", k=3, W=48, H=17, NB=3)
case(name="perf_k3_lr48", k=3, NB=256, H=48, W=48, iters=10, check=False)
case(name="perf_k5_lr48", k=5, NB=256, H=48, W=48, iters=10, check=False)
case(name="perf_k3_hr192", k=3, NB=32, H=192, W=192, iters=5, check=False)
case(name="perf_k5_hr192", k=5, NB=32, H=192, W=192, iters=5, check=False)
case(name="perf_k5_lr96", k=5, NB=148, H=96, W=96, iters=5, check=False)
case(name="perf_k5_hr192_nb256", k=5, NB=256, H=192, W=192, iters=5, check=False)   # the training step's HR shape
case(name="perf_k3_hr192_nb256", k=3, NB=256, H=192, W=192, iters=5, check=False)


def run_case(idx):
    import torch
    from sr100 import _lib as L
    cs = CASES[idx]
    lib = L.require_device()
    torch.manual_seed(idx)
    dev = "cuda"
    NB, H, W, k = cs["NB"], cs["H"], cs["W"], cs["k"]
    x = (torch.randn(NB, H, W, 128, device=dev) * 0.5).to(torch.bfloat16).contiguous()
    g = (torch.randn(NB, H, W, 128, device=dev) * 0.5).to(torch.bfloat16).contiguous()
    dw = torch.randn(k, k, 128, 128, device=dev)
    dw0 = dw.clone()
    ws = torch.empty(lib.sr_wgrad_workspace_bytes(), dtype=torch.uint8, device=dev)
    d = L.WgradDesc()
    d.x_bf16, d.g_bf16 = x.data_ptr(), g.data_ptr()
    d.NB, d.H, d.W, d.ksize = NB, H, W, k
    d.scale, d.accumulate = cs["scale"], cs["accumulate"]
    d.dw_hwio, d.workspace, d.workspace_bytes = dw.data_ptr(), ws.data_ptr(), ws.numel()
    plan = C.c_void_p()
    L.check(lib.sr_wgrad_plan_create(C.byref(d), C.byref(plan)))
    info = L.WgradPlanInfo()
    L.check(lib.sr_wgrad_plan_info(plan, C.byref(info)))
    L.check(lib.sr_wgrad_plan_run(plan, L.stream_ptr()))
    torch.cuda.synchronize()
    rec = dict(case=idx, **cs, grid=info.grid, smem=info.smem_bytes, seg_width=info.seg_width, nseg=info.nseg,
               ring=info.ring_rows, g_slots=info.g_slots, groups=info.tap_groups, rows_per_unit=info.rows_per_unit, images_per_row=info.images_per_row)
    if cs["check"]:
        xin = x.float().permute(0, 3, 1, 2).contiguous()
        gout = g.float().permute(0, 3, 1, 2).contiguous()
        want = torch.nn.grad.conv2d_weight(xin, (128, 128, k, k), gout, padding=k // 2)  # OIHW
        want = want.permute(2, 3, 1, 0).contiguous() * cs["scale"]                         # -> HWIO
        if cs["accumulate"]:
            want = want + dw0
        err = (dw - want).abs()
        ref = float(want.abs().max())
        rec["max_err"] = float(err.max())
        rec["max_ref"] = ref
        rec["per_tap_err"] = [round(float(err[t // k, t % k].max()), 4) for t in range(k * k)]
        rec["ok"] = bool(rec["max_err"] <= 2e-3 * max(ref, 1.0))
    if cs["iters"]:
        for _ in range(2):
            L.check(lib.sr_wgrad_plan_run(plan, L.stream_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(cs["iters"]):
            L.check(lib.sr_wgrad_plan_run(plan, L.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / cs["iters"]
        rec["ms"] = round(ms, 4)
        rec["tflops"] = round(info.flops / ms / 1e9, 1)
    lib.sr_wgrad_plan_destroy(plan)
    print(json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=None)
    ap.add_argument("--only", type=str, default=None)
    ap.add_argument("--out", type=str, default=os.path.join(ROOT, "gpurun_out", "probe_wgrad.jsonl"))
    a = ap.parse_args()
    if a.case is not None:
        try:
            run_case(a.case)
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(case=a.case, name=CASES[a.case]["name"], error=str(e)[-400:])), flush=True)
            sys.exit(3)
        return
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        for i, cs in enumerate(CASES):
            if a.only and a.only not in cs["name"]:
                continue
            cmd = [sys.executable, os.path.abspath(__file__), "--case", str(i)]
            try:
                p = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
                lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
                if not lines:
                    lines = [json.dumps(dict(case=i, name=cs["name"], rc=p.returncode, stderr=p.stderr[-600:],
                                             stdout=p.stdout[-300:]))]
            except subprocess.TimeoutExpired:
                lines = [json.dumps(dict(case=i, name=cs["name"], timeout=True))]
            for l in lines:
                f.write(l + "\n")
                print(l, flush=True)
            f.flush()


if __name__ == "__main__":
    main()
