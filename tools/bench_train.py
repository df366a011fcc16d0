"""Training-step benchmark (BASELINE.json configs[3]): 48x48 LR patches, global batch 256, forward + dgrad/wgrad
in bf16 on the tensor cores, NCCL gradient all-reduce, fused Adam.  Strong scaling: the global batch is split
across ranks.

    python tools/bench_train.py [--batch 256] [--size 48] [--steps 5] [--warmup 2]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py ...

Prints one JSON line on rank 0 (step ms = max over ranks, images/s, algorithmic TFLOP/s of the conv + wgrad
launches per GPU, share of the step spent in the all-reduce + Adam + repack tail)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=48)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--detail", action="store_true", help="per-launch CUDA-event breakdown of one step")
    ap.add_argument("--data", choices=["none", "device", "host"], default="none",
                    help="feed every step from a PNG dataset: 'device' = HBM-resident set (sr100.dataset), 'host' = "
                         "the reference's per-batch decode (img_utils.image_generator); 'none' = fixed tensors")
    ap.add_argument("--dataset-size", type=int, default=1024)
    a = ap.parse_args()
    import torch
    from sr100 import dist as D
    from sr100.engine import Engine
    from sr100.train import Trainer
    rank, local_rank, world = D.init_process_group()
    torch.cuda.set_device(local_rank if world > 1 else 0)
    lo, hi = D.shard_range(a.batch, rank, world)
    nb = hi - lo
    eng = Engine()
    tr = Trainer(eng)
    g = tr.graph(nb, a.size, a.size)
    gen = torch.Generator(device="cuda").manual_seed(7 + rank)
    g.x_in.copy_(torch.rand(g.x_in.shape, device="cuda", generator=gen))
    g.y_true.copy_(torch.rand(g.y_true.shape, device="cuda", generator=gen))
    feed = None
    if a.data != "none":
        import tempfile
        import time
        import numpy as np
        from PIL import Image
        import img_utils
        d = tempfile.mkdtemp(prefix="sr100_ds_") + "/"
        os.makedirs(d + "X")
        os.makedirs(d + "y")
        rng = np.random.default_rng(rank)
        for i in range(a.dataset_size):
            Image.fromarray(rng.integers(0, 256, size=(a.size, a.size, 3), dtype=np.uint8)).save(d + "X/%05d.png" % i)
            Image.fromarray(rng.integers(0, 256, size=(4 * a.size, 4 * a.size, 3), dtype=np.uint8)).save(d + "y/%05d.png" % i)
        t0 = time.time()
        feed = img_utils.image_generator(d, scale_factor=1, batch_size=nb, shuffle=True,
                                         device_resident=(a.data == "device"))
        first = next(feed)
        load_s = time.time() - t0

    def step():
        if feed is not None:
            x, y = next(feed)
            tr._load(g, x, y)
        tr.step_device(g)

    for _ in range(a.warmup):
        step()
    D.barrier()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t_fb = 0.0
    e0.record()
    for _ in range(a.steps):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if feed is not None:
            x, y = next(feed)
            tr._load(g, x, y)
        s0.record()
        tr.forward_backward_device(g)
        s1.record()
        tr.apply_gradients()
        s1.synchronize()
        t_fb += s0.elapsed_time(s1)
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    ms = D.max_over_ranks(e0.elapsed_time(e1) / a.steps)
    if feed is not None:      # host decode does not show on the device clock: wall time of the same loop
        import time
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
        ms = D.max_over_ranks((time.time() - t0) * 1e3 / a.steps)
    fb_ms = t_fb / a.steps
    loss = float(g.loss_sum.item()) / g.n_local
    rec = dict(metric="train_step", n_gpus=world, global_batch=a.batch, per_gpu_batch=nb, lr_size=a.size,
               steps=a.steps, ms_per_step=round(ms, 3), images_per_s=round(a.batch / ms * 1e3, 1),
               fwd_bwd_ms=round(fb_ms, 3), update_ms=round(ms - fb_ms, 3),
               algorithmic_tflop_per_step_per_gpu=round(tr.step_flops(g) / 1e12, 3),
               tflops_per_gpu=round(tr.step_flops(g) / (ms * 1e-3) / 1e12, 1),
               loss=loss, sequencer=eng.sequencer,
               mem_gb=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2), data=a.data)
    if feed is not None:
        rec["dataset_pairs"] = a.dataset_size
        rec["first_batch_s"] = round(load_s, 2)
    if hasattr(g, "fwd_flops"):
        rec.update(fwd_tflop=round(g.fwd_flops / 1e12, 3), bwd_tflop=round(g.bwd_flops / 1e12, 3))
    if a.detail and rank == 0:
        assert eng.sequencer == "python", "--detail times the Python launch lists: run with SR100_PY_SEQUENCE=1"
        st = __import__("sr100._lib", fromlist=["x"]).stream_ptr()
        parts = {}
        for label, lst in (("fwd", g.fwd), ("bwd", g.bwd)):
            evs = []
            for f in lst:
                x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                x0.record()
                f(st)
                x1.record()
                evs.append((x0, x1, getattr(f, "name", "other")))
            torch.cuda.synchronize()
            for x0, x1, kind in evs:
                parts[label + ":" + kind] = parts.get(label + ":" + kind, 0.0) + x0.elapsed_time(x1)
        rec["detail_ms"] = {k: round(v, 3) for k, v in parts.items()}
    if rank == 0:
        print(json.dumps(rec), flush=True)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
