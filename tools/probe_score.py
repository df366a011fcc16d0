"""Scoring kernel alone: time of one 1356x2040 pair (L2 evicted between launches) and agreement with the oracle
on a few pairs (noise levels from SR-like to unrelated images).  Dev tool for profiles/.

    python tools/probe_score.py [--out gpurun_out/probe_score.jsonl] [--once]     (--once: two launches, for ncu -k)
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
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe_score.jsonl"))
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    import numpy as np
    import torch
    from scipy.ndimage import uniform_filter
    from sr100 import _lib as L
    import scorpath
    lib = L.require_device()
    dev = "cuda"
    st = L.stream_ptr
    rng = np.random.default_rng(0)
    H, W = 1356, 2040
    gt = uniform_filter(rng.integers(0, 256, size=(H, W, 3)).astype(np.float32), size=(3, 3, 1)).astype(np.uint8)
    sr = np.clip(gt.astype(int) + rng.integers(-8, 9, size=gt.shape), 0, 255).astype(np.uint8)
    a8, b8 = torch.from_numpy(gt).to(dev), torch.from_numpy(sr).to(dev)
    res = torch.zeros(128, dtype=torch.uint8, device=dev)

    def run():
        res.zero_()
        L.check(lib.sr_score_pair_u8(L.ptr(a8), L.ptr(b8), H, W, 10, L.ptr(res), st()))
    run(); run()
    torch.cuda.synchronize()
    if a.once:
        return
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    tot = 0.0
    for _ in range(a.iters):
        flush.fill_(1)
        flush_rd.sum()
        res.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.sr_score_pair_u8(L.ptr(a8), L.ptr(b8), H, W, 10, L.ptr(res), st()))
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    recs = [dict(kernel="score_pair_kernel", shape=[H, W], ms=round(tot / a.iters, 4),
                 gbs=round(2 * H * W * 3 / (tot / a.iters) / 1e6, 1))]
    print(json.dumps(recs[0]), flush=True)
    # the config-3 batch: 64 pairs of 1356x2040 in two launches of 32 (1.06 GB of input: nothing stays in L2)
    from sr100 import ops
    import ctypes as C
    n = 64
    big_a = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device=dev)
    big_b = torch.clamp(big_a.to(torch.int16) + torch.randint(-8, 9, big_a.shape, dtype=torch.int16, device=dev), 0, 255).to(torch.uint8)
    items = (L.ScoreItem * n)()
    for i in range(n):
        items[i].a, items[i].b, items[i].h, items[i].w = L.ptr(big_a[i]), L.ptr(big_b[i]), H, W
    resn = torch.zeros(n * C.sizeof(L.ScoreResult), dtype=torch.uint8, device=dev)
    tot = 0.0
    for _ in range(5):
        resn.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.sr_score_batch_u8(C.cast(items, C.c_void_p), n, 10, L.ptr(resn), st()))
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    rec = dict(kernel="score_pair_kernel (64 pairs, sr_score_batch_u8)", shape=[n, H, W], ms=round(tot / 5, 4),
               ms_per_pair=round(tot / 5 / n, 4), gbs=round(n * 2 * H * W * 3 / (tot / 5) / 1e6, 1))
    recs.append(rec)
    print(json.dumps(rec), flush=True)
    del big_a, big_b
    # agreement with the oracle: SR-like noise, heavy noise, unrelated images; small and large
    from oracle import scoring as osc
    for (h, w, noise, smooth) in ((40, 37, 6, 5), (64, 48, 6, 5), (300, 200, 6, 5), (300, 200, 40, 1),
                                  (300, 200, 128, 1), (64, 64, 255, 1), (1356, 2040, 8, 3)):
        g = uniform_filter(rng.integers(0, 256, size=(h, w, 3)).astype(np.float32), size=(smooth, smooth, 1)).astype(np.uint8)
        s = np.clip(g.astype(int) + rng.integers(-noise, noise + 1, size=g.shape), 0, 255).astype(np.uint8)
        if noise == 255:
            s = rng.integers(0, 256, size=g.shape).astype(np.uint8)
        p, srgb, sy = scorpath.score_pair(g, s, 10)
        p2, srgb2, sy2 = scorpath.score_pair(g, s, 10)
        wp, wrgb, wy = osc.score_pair(g, s, 10)
        rec = dict(shape=[h, w], noise=noise, psnr_err=abs(p - wp), ssim_rgb_err=abs(srgb - wrgb), ssim_y_err=abs(sy - wy),
                   ssim_y=wy, reproducible=bool(p == p2 and srgb == srgb2 and sy == sy2))
        recs.append(rec)
        print(json.dumps(rec), flush=True)
    with open(a.out, "w") as f:
        for r in recs:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
