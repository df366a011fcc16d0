"""Where the time of BASELINE config 1 (one 128x128 patch -> 512x512, model.predict) goes: CUDA-graph replay time of
the whole forward, and per-launch CUDA-event times of the same launches run eagerly (one at a time, synchronised, so
the host launch rate does not pollute them), grouped by kernel kind with the plan geometry.  Dev tool for profiles/.

    python tools/probe_config1.py [--h 128 --w 128 --nb 1]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h", type=int, default=128)
    ap.add_argument("--w", type=int, default=128)
    ap.add_argument("--nb", type=int, default=1)
    ap.add_argument("--iters", type=int, default=200)
    a = ap.parse_args()
    import torch
    from sr100 import _lib as L
    from sr100.engine import Engine, _Plan, glorot_uniform_weights
    eng = Engine(glorot_uniform_weights(seed=1234), sequencer="python")     # per-launch access needs the Python launch lists
    x = torch.rand(a.nb, a.h, a.w, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    for _ in range(5):
        eng.forward_device(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        eng.forward_device(x)
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1) / a.iters
    flops = eng.last_flops()
    st = L.stream_ptr()
    groups = {}
    for stg in eng.last_stages:
        res = "hr" if hasattr(stg, "eh") else "lr"
        for step in stg.steps:
            plan = getattr(step, "__self__", None)
            best = 1e9
            for _ in range(5):
                torch.cuda.synchronize()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                step(st)
                s1.record()
                torch.cuda.synchronize()
                best = min(best, s0.elapsed_time(s1))
            if isinstance(plan, _Plan):
                i = plan.info
                key = "%s conv flops=%.3g grid=%d tiles=%d nseg=%d segw=%d T=%d stages=%d" % (
                    res, plan.flops, i.grid, i.total_tiles, i.nseg, i.seg_width, i.tile_positions, i.num_wstages)
            else:
                key = res + " other"
            g = groups.setdefault(key, dict(n=0, ms=0.0, flops=0.0))
            g["n"] += 1
            g["ms"] += best
            g["flops"] += plan.flops if isinstance(plan, _Plan) else 0.0
    rows = []
    for k, g in groups.items():
        rows.append(dict(kind=k, launches=g["n"], ms=round(g["ms"], 4), us_per_launch=round(g["ms"] / g["n"] * 1e3, 2),
                         tflops=round(g["flops"] / max(g["ms"], 1e-9) / 1e9, 1)))
    eager_sum = sum(g["ms"] for g in groups.values())
    print(json.dumps(dict(shape=[a.nb, a.h, a.w], graph_replay_ms=round(total_ms, 4),
                          tflops=round(flops / total_ms / 1e9, 1), sum_of_isolated_launch_ms=round(eager_sum, 4),
                          launches=sum(g["n"] for g in groups.values()), rows=rows), indent=1))


if __name__ == "__main__":
    main()
