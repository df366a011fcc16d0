"""GPU probe for the tcgen05 conv kernel: runs each (mode, geometry) case in its own subprocess under
a timeout, cross-checks against the CUDA-core direct conv, and reports TFLOP/s.

    python tools/probe_conv.py            # all cases -> gpurun_out/probe.jsonl
    python tools/probe_conv.py --case N   # one case (used by the driver loop)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))

CASES = []


def case(**kw):
    d = dict(NB=2, H=20, W=96, ks=(3,), cout=128, mode=0, nacc=4, res=None, relu=0, alpha=1.0, beta=0.0,
             iters=0, check=True, outs="both", pair=0, prec="bf16")
    d.update(kw)
    CASES.append(d)


for mode in (0, 1):
    case(name="k1_w128_aligned", ks=(1,), W=128, H=12, mode=mode)
    case(name="k1_w96", ks=(1,), mode=mode)
    case(name="k3_w96", ks=(3,), mode=mode, relu=1)
    case(name="k5_w96", ks=(5,), mode=mode)
    case(name="k5k3_dual_res", ks=(5, 3), mode=mode, res="f32", alpha=0.1, beta=0.9)
    case(name="k3_res_bf16", ks=(3,), mode=mode, res="bf16", alpha=0.1, beta=1.0)
    case(name="k3_tail3", ks=(3,), cout=3, mode=mode, relu=1)
    case(name="k5_nacc2", ks=(5,), mode=mode, nacc=2)
    case(name="k3_w50_h33", ks=(3,), W=50, H=33, NB=3, mode=mode)
    case(name="k5_w384", ks=(5,), W=384, H=24, NB=1, mode=mode)
    case(name="perf_k5_lr", ks=(5,), NB=32, H=96, W=96, mode=mode, iters=40, check=False)
    case(name="perf_k3_lr", ks=(3,), NB=32, H=96, W=96, mode=mode, iters=40, check=False)
    case(name="perf_k5_hr", ks=(5,), NB=2, H=384, W=384, mode=mode, iters=20, check=False)
    case(name="perf_k5_lr_nacc2", ks=(5,), NB=32, H=96, W=96, mode=mode, nacc=2, iters=40, check=False)
    case(name="perf_k5k3_hr", ks=(5, 3), NB=2, H=384, W=384, mode=mode, iters=20, check=False,
         res="f32", alpha=0.1, beta=0.9)


# many-wave cases (>= 16 rounds of 148 CTAs) so the persistent scheduler's tail does not dominate
for nacc in (4, 2):
    sfx = "" if nacc == 4 else "_nacc2"
    case(name="big_k5_lr_relu_bf16" + sfx, ks=(5,), NB=148, H=96, W=96, relu=1, outs="bf16", iters=10, check=False, nacc=nacc)
    case(name="big_k3_lr_relu_bf16" + sfx, ks=(3,), NB=148, H=96, W=96, relu=1, outs="bf16", iters=20, check=False, nacc=nacc)
    case(name="big_k5k3_lr_end" + sfx, ks=(5, 3), NB=148, H=96, W=96, res="f32", alpha=0.1, beta=0.9, iters=10,
         check=False, nacc=nacc)
    case(name="big_k5_hr_relu_bf16" + sfx, ks=(5,), NB=8, H=384, W=384, relu=1, outs="bf16", iters=10, check=False, nacc=nacc)
    case(name="big_k5k3_hr_end_bf16" + sfx, ks=(5, 3), NB=8, H=384, W=384, res="bf16", alpha=0.1, beta=0.9, outs="bf16",
         iters=10, check=False, nacc=nacc)


# CTA-pair (cta_group::2) kernel: correctness, then many-wave throughput
for nacc in (4, 2):
    sfx = "_nacc%d" % nacc
    case(name="pair_k3" + sfx, ks=(3,), NB=2, pair=1, nacc=nacc, relu=1)
    case(name="pair_k5_odd_nb" + sfx, ks=(5,), NB=3, pair=1, nacc=nacc)
    case(name="pair_k5k3_res" + sfx, ks=(5, 3), NB=4, H=33, W=50, pair=1, nacc=nacc, res="f32", alpha=0.1, beta=0.9)
    case(name="pair_k5_w384" + sfx, ks=(5,), NB=2, H=24, W=384, pair=1, nacc=nacc)
for nacc in (4, 2):
    sfx = "_nacc%d" % nacc
    case(name="pbig_k5_lr_relu_bf16" + sfx, ks=(5,), NB=148, H=96, W=96, relu=1, outs="bf16", iters=10, check=False, nacc=nacc, pair=1)
    case(name="pbig_k3_lr_relu_bf16" + sfx, ks=(3,), NB=148, H=96, W=96, relu=1, outs="bf16", iters=20, check=False, nacc=nacc, pair=1)
    case(name="pbig_k5k3_lr_end" + sfx, ks=(5, 3), NB=148, H=96, W=96, res="f32", alpha=0.1, beta=0.9, iters=10,
         check=False, nacc=nacc, pair=1)
    case(name="pbig_k5_hr_relu_bf16" + sfx, ks=(5,), NB=8, H=384, W=384, relu=1, outs="bf16", iters=10, check=False, nacc=nacc, pair=1)
    case(name="pbig_k5k3_hr_end_bf16" + sfx, ks=(5, 3), NB=8, H=384, W=384, res="bf16", alpha=0.1, beta=0.9, outs="bf16",
         iters=10, check=False, nacc=nacc, pair=1)


# tf32 option (sr_conv_desc.precision = 1): many-wave throughput of the CTA-pair kernel (correctness: tests/test_gpu_tf32.py)
case(name="tf32big_k5_lr_relu", ks=(5,), NB=148, H=96, W=96, relu=1, iters=10, check=False, nacc=2, pair=1, prec="tf32")
case(name="tf32big_k3_lr_relu", ks=(3,), NB=148, H=96, W=96, relu=1, iters=20, check=False, nacc=2, pair=1, prec="tf32")
case(name="tf32big_k5k3_lr_end", ks=(5, 3), NB=148, H=96, W=96, res="f32", alpha=0.1, beta=0.9, iters=10, check=False,
     nacc=2, pair=1, prec="tf32", outs="both")
case(name="tf32big_k5_hr_relu", ks=(5,), NB=8, H=384, W=384, relu=1, iters=10, check=False, nacc=2, pair=1, prec="tf32")
case(name="tf32big_k5k3_hr_end", ks=(5, 3), NB=8, H=384, W=384, res="f32", alpha=0.1, beta=0.9, iters=10, check=False,
     nacc=2, pair=1, prec="tf32", outs="both")


def run_case(idx):
    import torch
    from sr100 import _lib as L
    cs = CASES[idx]
    lib = L.require_device()
    torch.manual_seed(idx)
    dev = "cuda"
    NB, H, W, cout = cs["NB"], cs["H"], cs["W"], cs["cout"]
    ins, ws, packed = [], [], []
    tf32 = cs["prec"] == "tf32"
    for k in cs["ks"]:
        x = torch.randn(NB, H, W, 128, device=dev) * 0.5
        w = (torch.randn(k, k, 128, cout, device=dev) / (k * k * 128) ** 0.5).contiguous()
        if tf32:
            L.check(lib.sr_round_tf32(L.ptr(x), x.numel(), L.ptr(x), L.stream_ptr()))
            pk = torch.empty(lib.sr_packed_weight_bytes_tf32(k, cout), dtype=torch.uint8, device=dev)
            L.check(lib.sr_pack_conv_weights_tf32(L.ptr(w), k, cout, L.ptr(pk), L.stream_ptr()))
        else:
            x = x.to(torch.bfloat16).contiguous()
            pk = torch.empty(lib.sr_packed_weight_bytes(k, cout), dtype=torch.uint8, device=dev)
            L.check(lib.sr_pack_conv_weights(L.ptr(w), k, cout, 0, L.ptr(pk), L.stream_ptr()))
        ins.append(x); ws.append(w); packed.append(pk)
    bias = torch.randn(cout, device=dev) * 0.1
    res = None
    if cs["res"] == "f32":
        res = torch.randn(NB, H, W, cout, device=dev)
    elif cs["res"] == "bf16":
        res = torch.randn(NB, H, W, cout, device=dev).to(torch.bfloat16)
    out_bf16 = torch.zeros(NB, H, W, cout, device=dev, dtype=torch.bfloat16)
    out_f32 = torch.zeros(NB, H, W, cout, device=dev)
    d = L.ConvDesc()
    d.nsrc = len(cs["ks"])
    for s, k in enumerate(cs["ks"]):
        d.in_[s] = ins[s].data_ptr()
        d.wpacked[s] = packed[s].data_ptr()
        d.ksize[s] = k
    d.NB, d.H, d.W, d.cin, d.cout = NB, H, W, 128, cout
    d.bias = bias.data_ptr()
    d.alpha, d.beta, d.relu = cs["alpha"], cs["beta"], cs["relu"]
    if cs["res"] == "f32":
        d.res_f32 = res.data_ptr()
    elif cs["res"] == "bf16":
        d.res_bf16 = res.data_ptr()
    if tf32:      # operand copy (tf32-rounded fp32) always, the unrounded stream for the block-end launches
        out_t32 = torch.zeros(NB, H, W, cout, device=dev)
        d.precision, d.out_tf32 = 1, out_t32.data_ptr()
        if cs["res"]:
            d.out_f32 = out_f32.data_ptr()
    else:
        d.out_bf16 = out_bf16.data_ptr()
        if cs["outs"] == "both":
            d.out_f32 = out_f32.data_ptr()
    d.a_mode, d.nacc, d.pair = cs["mode"], cs["nacc"], cs["pair"]
    plan = C.c_void_p()
    L.check(lib.sr_conv_plan_create(C.byref(d), C.byref(plan)))
    info = L.ConvPlanInfo()
    L.check(lib.sr_conv_plan_info(plan, C.byref(info)))
    L.check(lib.sr_conv_plan_run(plan, L.stream_ptr()))
    torch.cuda.synchronize()
    rec = dict(case=idx, **{k: v for k, v in cs.items()}, seg_width=info.seg_width, nseg=info.nseg,
               strip_rows=info.strip_rows, wstages=info.num_wstages, tiles=info.total_tiles,
               smem=info.smem_bytes, mma_eff=round(info.mma_efficiency, 4))
    if cs["check"]:
        acc = torch.zeros(NB, H, W, cout, device=dev)
        for s, k in enumerate(cs["ks"]):
            tmp = torch.empty(NB, H, W, cout, device=dev)
            L.check(lib.sr_conv2d_direct(L.ptr(ins[s]), 1, L.ptr(ws[s]), 1, None, NB, H, W, 128, cout, k,
                                         1, 0, 0, 0, L.ptr(tmp), L.stream_ptr()))
            acc += tmp
            # independent check of the direct kernel against cuDNN (dev tool only)
            ref = torch.nn.functional.conv2d(ins[s].float().permute(0, 3, 1, 2),
                                             ws[s].to(torch.bfloat16).float().permute(3, 2, 0, 1),
                                             padding=k // 2).permute(0, 2, 3, 1)
            rec["direct_vs_cudnn_%d" % s] = float((tmp - ref).abs().max())
        want = cs["alpha"] * (acc + bias)
        if res is not None:
            want = want + cs["beta"] * res.float()
        if cs["relu"]:
            want = want.clamp_min(0)
        err = (out_f32 - want).abs()
        rec["max_err_f32"] = float(err.max())
        rec["max_ref"] = float(want.abs().max())
        rec["max_err_bf16"] = float((out_bf16.float() - want).abs().max())
        if rec["max_err_f32"] > 1e-2:
            bad = (err > 1e-2).nonzero()
            rec["n_bad"] = int(bad.shape[0])
            rec["first_bad"] = bad[:6].tolist()
            rec["frac_bad"] = float(bad.shape[0]) / err.numel()
    if cs["iters"]:
        for _ in range(3):
            L.check(lib.sr_conv_plan_run(plan, L.stream_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(cs["iters"]):
            L.check(lib.sr_conv_plan_run(plan, L.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / cs["iters"]
        rec["ms"] = round(ms, 4)
        rec["tflops"] = round(info.flops / ms / 1e9, 1)
    lib.sr_conv_plan_destroy(plan)
    print(json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--from-case", type=int, default=None, help="run cases i.. in this process")
    ap.add_argument("--only", type=str, default=None, help="substring filter on case names")
    ap.add_argument("--out", type=str, default=os.path.join(ROOT, "gpurun_out", "probe.jsonl"))
    a = ap.parse_args()
    sel = [i for i, cs in enumerate(CASES) if not a.only or a.only in cs["name"]]
    if a.from_case is not None:
        # child: run sequentially; a CUDA fault poisons the context, so stop at the first failure
        for i in sel:
            if i < a.from_case:
                continue
            print("BEGIN %d" % i, flush=True)
            try:
                run_case(i)
            except Exception as e:  # noqa: BLE001
                print(json.dumps(dict(case=i, name=CASES[i]["name"], mode=CASES[i]["mode"],
                                      error=str(e)[-400:])), flush=True)
                sys.exit(3)
        return
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    nxt = 0
    with open(a.out, "w") as f:
        while nxt is not None and nxt < len(CASES):
            cmd = [sys.executable, os.path.abspath(__file__), "--from-case", str(nxt)]
            if a.only:
                cmd += ["--only", a.only]
            try:
                p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
                out, timed_out = p.stdout, False
            except subprocess.TimeoutExpired as e:
                out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                timed_out = True
            last_begin = None
            for l in out.splitlines():
                if l.startswith("BEGIN "):
                    last_begin = int(l.split()[1])
                elif l.startswith("{"):
                    f.write(l + "\n")
                    print(l, flush=True)
            f.flush()
            if timed_out or p.returncode != 0:
                if timed_out:
                    rec = dict(case=last_begin, name=CASES[last_begin]["name"] if last_begin is not None else None,
                               timeout=True)
                    f.write(json.dumps(rec) + "\n")
                    print(json.dumps(rec), flush=True)
                elif p.returncode != 3:
                    rec = dict(case=last_begin, rc=p.returncode, stderr=p.stderr[-500:])
                    f.write(json.dumps(rec) + "\n")
                    print(json.dumps(rec), flush=True)
                nxt = None if last_begin is None else last_begin + 1
            else:
                nxt = None


if __name__ == "__main__":
    main()
