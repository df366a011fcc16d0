"""Throughput of the other BASELINE.json configs (bench.py measures configs[1]); one JSON line per config on rank 0.

    python tools/bench_configs.py --config 1      # one 128x128 patch -> 512x512 through model.predict (latency)
    python tools/bench_configs.py --config 3      # 64 x (339x510) images, tiled, images round-robin over ranks
    python tools/bench_configs.py --config 5      # one 1080x1920 image, its tiles sharded over ranks, rank 0 stitches
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs.py --config 3|5

Config 5 under N > 1 also checks that the sharded result is bit-identical to the single-rank one."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def synth(rng, h, w):
    import numpy as np
    from scipy.ndimage import uniform_filter
    img = rng.integers(0, 256, size=(h + 4, w + 4, 3)).astype(np.float32)
    return uniform_filter(img, size=(5, 5, 1))[2:-2, 2:-2].astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import numpy as np
    import torch
    from sr100 import dist as D
    rank, local_rank, world = D.init_process_group()
    torch.cuda.set_device(local_rank if world > 1 else 0)
    import models
    m = models.DifvdsrDouble(1)
    model = m.create_model(96, 96)
    eng = model.engine
    from sr100.engine import glorot_uniform_weights
    eng.set_weights_dict(glorot_uniform_weights(seed=1234))     # every rank holds the SAME weight replica

    def timed(fn, iters, warmup):
        for _ in range(warmup):
            fn()
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        D.barrier()
        return D.max_over_ranks(e0.elapsed_time(e1) / iters)

    rec = dict(config=a.config, n_gpus=world, iters=a.iters)
    if a.config == 1:
        x = torch.rand(1, 128, 128, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        ms = timed(lambda: eng.forward_device(x), max(a.iters, 100), max(a.warmup, 20))
        flops = eng.last_flops()
        rec.update(workload="1 x 128x128 -> 512x512, model.predict path (CUDA-graph replay)", ms=round(ms, 4),
                   mp_per_s=round(0.262144 / ms * 1e3, 2), tflops=round(flops / ms / 1e9, 1))
    elif a.config == 3:
        rng = np.random.default_rng(100)
        imgs = [synth(rng, 339, 510) for _ in range(64)]
        mine = [torch.from_numpy(imgs[i]).cuda() for i in D.shard_round_robin(64, rank, world)]
        ms = timed(lambda: eng.upscale_images_device(mine), a.iters, a.warmup)
        flops = eng.last_flops()
        mp = 64 * 16 * 339 * 510 / 1e6
        rec.update(workload="64 x (339x510) -> 1356x2040, reference tiling with dead-work elimination, images "
                            "round-robin over ranks, inputs resident", ms=round(ms, 3), mp_per_s=round(mp / ms * 1e3, 1),
                   tflops_per_gpu=round(flops / ms / 1e9, 1), images_per_gpu=len(mine))
    elif a.config == 5:
        rng = np.random.default_rng(9)
        img = torch.from_numpy(synth(rng, 1080, 1920)).cuda()
        ms = timed(lambda: eng.upscale_image_sharded(img), a.iters, a.warmup)
        flops = eng.last_flops()
        out = eng.upscale_image_sharded(img)
        same = None
        if world > 1 and rank == 0:
            ref = eng.upscale_images_device([img])[0]
            same = bool(torch.equal(ref, out))
        mp = 16 * 1080 * 1920 / 1e6
        rec.update(workload="1 x (1080x1920) -> 4320x7680, 510 live tiles of 558 sharded over ranks, gather + stitch on "
                            "rank 0", ms=round(ms, 3), mp_per_s=round(mp / ms * 1e3, 1),
                   tflops_per_gpu=round(flops / ms / 1e9, 1), sharded_equals_single_rank=same)
    else:
        raise SystemExit("config must be 1, 3 or 5")
    if rank == 0:
        print(json.dumps(rec), flush=True)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
