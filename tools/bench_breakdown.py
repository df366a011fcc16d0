"""Per-launch breakdown of one forward of the headline bench workload (set5_x4_tiled): for every stage (LR stage,
each HR shape class) the tiles, executed conv FLOPs, summed launch time (CUDA events, eager launches) and TFLOP/s.

    python tools/bench_breakdown.py [--out gpurun_out/bench_breakdown.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_breakdown.json"))
    a = ap.parse_args()
    import torch
    import bench
    import models
    from sr100 import _lib as L
    from sr100.engine import _Plan, glorot_uniform_weights
    m = models.DifvdsrDouble(1)
    model = m.create_model(96, 96)
    eng = model.engine
    eng.set_weights_dict(glorot_uniform_weights(seed=1234))
    dev_imgs = [torch.from_numpy(im).cuda() for im in bench.synth_images(100)]
    for _ in range(3):
        eng.upscale_images_device(dev_imgs)
    torch.cuda.synchronize()
    st = L.stream_ptr()
    rows = []
    for rep in range(3):
        rows = []
        for stg in eng.last_stages:
            evs = []
            for step in stg.steps:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step(st)
                e1.record()
                evs.append((e0, e1, step))
            torch.cuda.synchronize()
            conv = [(e0.elapsed_time(e1), s.__self__) for e0, e1, s in evs if isinstance(getattr(s, "__self__", None), _Plan)]
            other = sum(e0.elapsed_time(e1) for e0, e1, s in evs if not isinstance(getattr(s, "__self__", None), _Plan))
            ms = sum(t for t, _ in conv)
            fl = sum(p.flops for _, p in conv)
            name = "LR %dx%dx%d" % (stg.NB, stg.H, stg.W) if hasattr(stg, "x_in") else "HR %d x %dx%d" % (stg.n, stg.eh, stg.ew)
            rows.append(dict(stage=name, conv_launches=len(conv), conv_ms=round(ms, 3), tflop=round(fl / 1e12, 3),
                             tflops=round(fl / ms / 1e9, 1) if ms else None, other_ms=round(other, 3),
                             min_launch_tflops=round(min(p.flops / t / 1e9 for t, p in conv), 1),
                             max_launch_tflops=round(max(p.flops / t / 1e9 for t, p in conv), 1)))
    tot_ms = sum(r["conv_ms"] for r in rows)
    tot_fl = sum(r["tflop"] for r in rows)
    for r in rows:
        r["share_of_conv_time"] = round(r["conv_ms"] / tot_ms, 4)
        print(json.dumps(r))
    print(json.dumps(dict(total_conv_ms=round(tot_ms, 3), total_tflop=round(tot_fl, 2), tflops=round(tot_fl / tot_ms * 1e3, 1))))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
