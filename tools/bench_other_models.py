"""Throughput of the two older graphs (sr100.planenet): Difvdsr4 (256 ch, x4) and Difvdsr (192 ch, x1) on a batch of
96x96 tiles.  Prints one JSON line per model: ms per pass, output MP/s, conv TFLOP/s (algorithmic FLOPs of the real
C x C contraction; the 192-channel model executes 256-wide launches, so its executed rate is (256/192)^2 higher).

    python tools/bench_other_models.py [--tiles 32] [--size 96] [--iters 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=32)
    ap.add_argument("--size", type=int, default=96)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--train", action="store_true", help="also time one optimizer step (sr100.planetrain)")
    ap.add_argument("--train-tiles", type=int, default=32)
    ap.add_argument("--train-size", type=int, default=48)
    a = ap.parse_args()
    import torch
    from sr100.planenet import PlaneNet
    for arch in ("difvdsr4", "difvdsr"):
        eng = PlaneNet(arch)
        x = torch.rand(a.tiles, a.size, a.size, 3, device="cuda")
        net = eng.net(a.tiles, a.size, a.size)
        net.x_in.copy_(x)
        for _ in range(3):
            net.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            net.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        C = eng.C
        s = eng.scale
        px = a.tiles * a.size * a.size
        if arch == "difvdsr4":
            macs = px * (3 * C + 12 * 9 * C * C) + 4 * px * 40 * 9 * C * C + 16 * px * (12 * 9 * C * C + 9 * C * 3)
        else:
            macs = px * (27 * C + 128 * 9 * C * C + 9 * C * 3)
        print(json.dumps(dict(model=arch, tiles=a.tiles, size=a.size, ms=round(ms, 3),
                              out_mp_per_s=round(px * s * s / ms / 1e3, 2),
                              algorithmic_tflops=round(2 * macs / ms / 1e9, 1),
                              executed_tflops=round(net.conv_flops / ms / 1e9, 1),
                              mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))), flush=True)
        del net
        if a.train:
            from sr100.planetrain import PlaneTrainer
            tr = PlaneTrainer(eng)
            g = tr.graph(a.train_tiles, a.train_size, a.train_size)
            g.x_in.copy_(torch.rand(g.x_in.shape, device="cuda"))
            g.y_true.copy_(torch.rand(g.y_true.shape, device="cuda"))
            for _ in range(2):
                tr.step_device(g)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(a.iters):
                tr.step_device(g)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            print(json.dumps(dict(model=arch, train_step=True, tiles=a.train_tiles, size=a.train_size, ms=round(ms, 3),
                                  images_per_s=round(a.train_tiles / ms * 1e3, 1),
                                  executed_tflops=round(g.flops / ms / 1e9, 1), launches=len(g.fwd) + len(g.bwd),
                                  mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))), flush=True)
            del tr, g
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
