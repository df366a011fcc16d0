"""Benchmark of the x4 SR hot path (BASELINE.json metric: x4 output megapixels/sec at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--only headline,set5,train,config5,whole,config1]

Headline workload (config.workload = "batch64_339x510_x4_tiled", BASELINE.json configs[2]):
64 synthetic RGB images of 339x510 (seeds 100..163) -> 1356x2040 through the reference's tiling (96/64 patches,
54 tiles per image, 8-px crop stitch; models.py:184-415) and the 86-conv DifvdsrDouble stack with random-init
weights.  One step = one pass over all 64 images (177.0 output MP); the images are dealt round-robin to the N
ranks (one process per GPU, torchrun), every rank holds a weight replica, there is no data-path collective:
STRONG scaling, the total work is fixed.  value = 177.0 MP x K / max-over-ranks device time.

value: this rank's images resident in HBM when the timed region starts (CUDA events).  e2e: the same step through
the public API (models.DifvdsrDouble.upscale_arrays) with pinned HOST buffers: H2D of the uint8 inputs and D2H of
the uint8 outputs inside the timed region (wall clock around a device synchronize).

The same JSON line carries the other BASELINE configs as objects (each its own timed region, same rules):
  "set5"       configs[1]: five Set5-shaped images tiled + Y-PSNR / Y-SSIM / RGB-SSIM scoring (every rank its own copy)
  "train_step" configs[3]: 48x48 LR patches, GLOBAL batch 256 split over the ranks, forward + dgrad/wgrad on the
               tensor cores, then ONE fused reduce-scatter + Adam + all-gather kernel per rank over NVLink peer memory
               (csrc/exchange.cu; the ncclAllReduce + Adam variant is timed beside it) (strong scaling)
  "config5"    configs[4]: one 1080x1920 image, its 510 live tiles sharded over the ranks, rank 0 stitches
  "whole_image" configs[2] again in whole-image mode (model.predict of the 339x510 images, no tiling)
  "config1"    configs[0]: one 128x128 patch -> 512x512 through model.predict (latency)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "image-enhance-keras_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

SET5_SHAPES = [(512, 512), (288, 288), (256, 256), (280, 280), (344, 228)]   # (H, W) of the Set5 GT files
B_IMAGES, B_H, B_W = 64, 339, 510                                             # BASELINE configs[2]
METRIC = "x4_output_megapixels_per_sec"
UNIT = "MP/s"


def synth_image(seed, h, w):
    """uint8 [h,w,3]: uniform noise through a 5x5 box blur (not white noise), SURVEY.md 8d."""
    import numpy as np
    from scipy.ndimage import uniform_filter
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h + 4, w + 4, 3)).astype(np.float32)
    return uniform_filter(img, size=(5, 5, 1))[2:-2, 2:-2].astype(np.uint8)


def synth_set5(seed):
    return [synth_image(seed * 16 + i, h, w) for i, (h, w) in enumerate(SET5_SHAPES)]


def headline_mp():
    return B_IMAGES * 16 * B_H * B_W / 1e6


def headline_config():
    """The static description of the headline workload: identical in both arms (own and --impl reference)."""
    return {"workload": "batch64_339x510_x4_tiled", "baseline_config": 2, "images": B_IMAGES, "lr_shape": [B_H, B_W],
            "hr_shape": [4 * B_H, 4 * B_W], "patch": 96, "step": 64, "tiles_per_image": 54, "tiles": 54 * B_IMAGES,
            "output_mp_per_step": headline_mp(), "sharding": "images round-robin over ranks, no data-path collective",
            "weights": "glorot_uniform random init (seed 1234), the same replica on every rank",
            "l2": "every conv launch streams 0.36-2.9 GB of activations (> 126 MB L2); no flush needed"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # the sampler also sees idle moments at the edges: report the median of the upper half
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], src="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


def cpu_reference_rate(threads, budget_s=15.0, max_tiles=25):
    """Oracle (CPU restatement of the reference graph; Keras/TF are not installable offline) on a bounded sample of
    the headline workload: the first n of the 54 96x96 tiles of image 0 (339x510, the reference's literal tiling:
    every tile through the full network).  MP = n/54 of that image's 2.766 output MP."""
    import numpy as np
    import torch
    from oracle import model as om
    from oracle import tiling as ot
    torch.set_num_threads(threads)
    img = synth_image(100, B_H, B_W)
    canvas = ot.make_canvas(img, 96, 64)
    patches, counts = ot.extract_patches_step(canvas, (96, 96), 64)
    assert patches.shape[0] == 54
    x = patches.astype(np.float32) / 255.
    weights = om.init_weights(1234)
    m = om.DifvdsrDoubleOracle(weights)
    with torch.no_grad():
        t0 = time.time()
        m(torch.from_numpy(x[:1]))
        t1 = time.time() - t0
        n = int(max(1, min(max_tiles, budget_s / max(t1, 1e-3))))
        t0 = time.time()
        y = m(torch.from_numpy(x[:n]))
        dt = time.time() - t0
    assert y.shape[0] == n
    mp = (n / 54.0) * (16 * B_H * B_W / 1e6)
    return mp / dt, "first %d of the 54 96x96 tiles of one 339x510 image, torch CPU fp32 oracle, %.1f s" % (n, dt), dt


def run_reference(args):
    """The reference arm: the reference's CPU path for the same metric/config on the host cores.  Keras 2 / TF 1 cannot
    be installed offline, so it is the CPU restatement (oracle/model.py, kind 'port'); each step times a bounded sample
    (8 tiles) of the workload.  Rank 0 alone runs it -- ONE CPU process whatever N is, so the driver's ratio at N > 1
    compares N GPUs with the same single host."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    sample = ""
    for i in range(args.warmup + args.steps):
        v, sample, _ = cpu_reference_rate(threads, budget_s=8.0, max_tiles=8)
        if i >= args.warmup:
            vals.append(v)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": headline_mp() / value * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample + " per step (ms_per_step scales that rate to the full 64-image step); "
                                            "Keras/TensorFlow cannot be installed offline, so this is the CPU "
                                            "restatement of the reference graph (oracle/model.py); one host process "
                                            "whatever --gpus is"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="sr100")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only", type=str, default="headline,set5,train,config5,whole,config1",
                    help="comma list of parts to run (headline is needed for a valid line)")
    ap.add_argument("--profile", action="store_true", help="ncu pass: honour --warmup as given, skip the extras")
    ap.add_argument("--allow-dev-build", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from sr100 import dist as D
    from sr100 import _lib as L
    rank, local_rank, world = D.init_process_group()
    torch.cuda.set_device(local_rank if world > 1 else 0)
    import models
    import scorpath
    from sr100.engine import glorot_uniform_weights
    lib = L.require_device()
    if lib.sr_dev_switches() and not args.allow_dev_build:
        raise SystemExit("libsr100.so is a development build (-DSR_DEV_SWITCHES): its numbers are not bench values")
    parts = set(args.only.split(","))
    if args.profile:
        parts = {"headline"}
    warmup = args.warmup if args.profile else max(args.warmup, 3)
    steps = args.steps

    def pinned(a):
        """Host arrays of the end-to-end path live in page-locked memory (what a serving loop would reuse)."""
        t = torch.empty(a.shape, dtype=torch.uint8, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()

    def timed(fn, k, w):
        """W warm-up calls, then K calls between barrier + synchronize, CUDA events, max over ranks (seconds)."""
        for _ in range(w):
            fn()
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        D.barrier()
        return D.max_over_ranks(e0.elapsed_time(e1) / 1e3)

    def timed_wall(fn, k, w):
        """The end-to-end variant: wall clock brackets the host work too (events alone would miss the D2H waits)."""
        for _ in range(w):
            fn()
        D.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        D.barrier()
        return D.max_over_ranks(dt)

    m = models.DifvdsrDouble(1)
    model = m.create_model(96, 96)                      # random-init glorot weights (no trained weights offline)
    eng = model.engine
    eng.set_weights_dict(glorot_uniform_weights(seed=1234))   # the same replica on every rank
    peaks = load_peaks()
    line = {}

    # =============================================================== headline: configs[2], 64 x 339x510, sharded by image
    my_ids = D.shard_round_robin(B_IMAGES, rank, world)
    images = [pinned(synth_image(100 + i, B_H, B_W)) for i in my_ids]
    dev_imgs = [torch.from_numpy(im).cuda() for im in images]

    def step_resident(full=False):
        return eng.upscale_images_device(dev_imgs, full_canvas=full)

    def step_e2e():
        return m.upscale_arrays(images)

    sampler = ClockSampler(local_rank if world > 1 else 0)
    if rank == 0:
        sampler.start()
    t_res = timed(step_resident, steps, warmup)
    clocks = sampler.stop() if rank == 0 else None
    mp_step = headline_mp()
    value = mp_step * steps / t_res
    e2e_steps = steps if not args.profile else 1
    t_e2e = timed_wall(step_e2e, e2e_steps, 2 if not args.profile else 0)
    e2e = mp_step * e2e_steps / t_e2e

    # ---- roofline of the dominant kernel (conv_tc_pair_kernel): per-launch CUDA events over one more step
    # (sr_model_forward_timed: the same launches issued one by one between event pairs on the launching stream)
    step_resident()
    torch.cuda.synchronize()
    tiles_run, hr_shapes = eng.last_run_summary()
    recs = eng.timed_launches()                      # [(ms, flops)], flops > 0 only for tensor-core conv launches
    conv_ms = sum(ms for ms, fl in recs if fl > 0)
    n_conv = sum(1 for ms, fl in recs if fl > 0)
    total_ms = sum(ms for ms, _ in recs)
    evs = recs
    conv_flops = eng.last_flops()
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12
    rpeaks = peaks
    if eng.tf32:     # SR100_PRECISION=tf32: kind::tf32 MMAs run at half the bf16 rate; no measured tf32 peak exists
        rpeaks = {"sustained": peaks["sustained"] / 2, "burst": peaks["burst"] / 2,
                  "src": peaks["src"] + " (halved: tf32 = 0.5x bf16 tensor rate)"}
    # libsr100 kernels per step on this rank: the stage launches + one batched gather (+ one stitch per image when the
    # stitch is not fused into the tail convs' epilogues)
    fused_stitch = eng.sequencer == "c" and os.environ.get("SR100_FUSED_STITCH", "1") != "0"
    launches_per_step = len(evs) + 1 + (0 if fused_stitch else len(dev_imgs))
    traffic, traffic_src = None, None
    prof_dir = os.path.join(ROOT, "profiles")
    for cand in sorted(os.listdir(prof_dir), reverse=True) if os.path.isdir(prof_dir) and not eng.tf32 else []:
        if cand.endswith("_ncu_step_summary.json"):
            try:
                traffic = json.load(open(os.path.join(prof_dir, cand)))["conv_traffic_bytes_per_launch"]
                traffic_src = "profiles/" + cand + " (ncu dram__bytes_read+write, mean over the conv launches of one forward)"
            except (OSError, KeyError, ValueError):
                pass
            break
    roofline = {"bound": "tensor", "kernel": "conv_tc_pair_kernel", "achieved": round(achieved, 1),
                "peak": rpeaks["sustained"], "unit": "TFLOP/s", "frac": round(achieved / rpeaks["sustained"], 4),
                "peak_source": rpeaks["src"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                "frac_of_burst_peak": round(achieved / rpeaks["burst"], 4), "traffic": traffic,
                "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                "launches": n_conv, "avg_launch_ms": round(conv_ms / n_conv, 4),
                "share_of_forward": round(conv_ms / total_ms, 4),
                "executed_flops_per_step_this_rank": conv_flops,
                "reference_tiling_flops_per_step": 54.0 * B_IMAGES * eng.conv_flops(1, 96, 96)}
    # the reference's literal tile set (all 54 tiles per image, full 384x384 HR stage): same output pixels
    value_full = None
    if not args.profile:
        k_full = 2
        t_full = timed(lambda: step_resident(True), k_full, 1)
        value_full = mp_step * k_full / t_full
    h2d = sum(im.nbytes for im in images)
    d2h = sum(16 * im.nbytes for im in images)
    h2d, d2h = int(D.sum_over_ranks(h2d)), int(D.sum_over_ranks(d2h))
    launches_total = int(D.sum_over_ranks(launches_per_step)) * steps
    eng.release()

    # =============================================================== configs[1]: Set5 tiled + scoring (every rank its own copy)
    if "set5" in parts:
        s_imgs = [pinned(im) for im in synth_set5(100 + rank)]
        s_dev = [torch.from_numpy(im).cuda() for im in s_imgs]
        rng = np.random.default_rng(7 + rank)
        gts = [pinned(rng.integers(0, 256, size=(4 * h, 4 * w, 3)).astype(np.uint8)) for h, w in SET5_SHAPES]
        dev_gts = [torch.from_numpy(g).cuda() for g in gts]
        score_buf = torch.zeros(5, 112, dtype=torch.uint8, device="cuda")   # five sr_score_result (112 bytes each)

        import ctypes as C
        items = (L.ScoreItem * 5)()

        def set5_resident():
            canv = eng.upscale_images_device(s_dev)
            score_buf.zero_()
            for i, (c, g) in enumerate(zip(canv, dev_gts)):
                items[i].a, items[i].b, items[i].h, items[i].w = L.ptr(c), L.ptr(g), g.shape[0], g.shape[1]
            L.check(lib.sr_score_batch_u8(C.cast(items, C.c_void_p), 5, 10, L.ptr(score_buf), L.stream_ptr()))   # one launch

        def set5_e2e():
            outs = m.upscale_arrays(s_imgs)
            return scorpath.score_pairs(list(zip(gts, outs)), 10)

        mp5 = sum(16 * h * w for h, w in SET5_SHAPES) / 1e6
        k5 = max(steps, 10)
        t5 = timed(set5_resident, k5, 3)
        t5e = timed_wall(set5_e2e, k5, 2)
        line["set5"] = {"workload": "set5_x4_tiled (BASELINE configs[1]): 5 images, 186 tiles (154 run), gather + stack + "
                                    "stitch + Y-PSNR/Y-SSIM/RGB-SSIM; every rank runs its own copy (replicas)",
                        "value_per_gpu": round(mp5 * k5 / t5, 3), "e2e_per_gpu": round(mp5 * k5 / t5e, 3), "unit": UNIT,
                        "ms_per_step": round(t5 / k5 * 1e3, 3), "steps": k5}
        eng.release()

    # =============================================================== configs[4]: one 1080x1920 image, tiles sharded over ranks
    if "config5" in parts:
        big = torch.from_numpy(synth_image(9, 1080, 1920)).cuda()
        k = max(3, min(steps, 5))
        t = timed(lambda: eng.upscale_image_sharded(big), k, 3)
        mpb = 16 * 1080 * 1920 / 1e6
        line["config5"] = {"workload": "1 x (1080x1920) -> 4320x7680 (BASELINE configs[4]): 510 live tiles of 558 sharded "
                                       "over the ranks in contiguous column-major ranges, " + eng.sharded_gather_description, "value": round(mpb * k / t, 2), "unit": UNIT,
                           "ms_per_step": round(t / k * 1e3, 3), "steps": k, "scaling": "strong"}
        del big
        eng.release()

    # =============================================================== configs[2] secondary: whole-image mode
    if "whole" in parts:
        xw = torch.stack([d.float() for d in dev_imgs]).div_(255.0) if dev_imgs else None
        outw = torch.empty(len(dev_imgs), 4 * B_H, 4 * B_W, 3, device="cuda") if dev_imgs else None
        kw = max(2, min(steps, 3))
        tw = timed(lambda: eng.forward_device(xw, out=outw) if xw is not None else None, kw, 3)
        fw = D.sum_over_ranks(eng.conv_flops(len(dev_imgs), B_H, B_W))
        line["whole_image"] = {"workload": "the same 64 x 339x510 images as whole images through model.predict (the "
                                           "reference's upVideo path, models.py:165-182: no tiling, no 2.5x tile "
                                           "redundancy), images round-robin over the ranks, fp32 [0,1] in HBM -> fp32 x4 out",
                               "value": round(mp_step * kw / tw, 3), "unit": UNIT, "ms_per_step": round(tw / kw * 1e3, 3),
                               "steps": kw, "tflops": round(fw * kw / tw / 1e12, 1),
                               "frac_of_sustained_peak": round(fw * kw / tw / 1e12 / peaks["sustained"] / world, 4),
                               "scaling": "strong"}
        del xw, outw
        eng.release()

    # =============================================================== configs[0]: one 128x128 patch (latency)
    if "config1" in parts:
        x1 = torch.rand(1, 128, 128, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        t = timed(lambda: eng.forward_device(x1), 100, 20)
        f1 = eng.last_flops()
        line["config1"] = {"workload": "1 x 128x128 -> 512x512 through model.predict (BASELINE configs[0]), CUDA-graph replay; "
                                       "every rank runs its own copy", "ms": round(t / 100 * 1e3, 4),
                           "value_per_gpu": round(0.262144 * 100 / t, 2), "unit": UNIT,
                           "tflops": round(f1 * 100 / t / 1e12, 1),
                           "frac_of_sustained_peak": round(f1 * 100 / t / 1e12 / peaks["sustained"], 4)}
        eng.release()

    # =============================================================== configs[3]: training step, global batch 256
    if "train" in parts and not eng.tf32:
        from sr100.train import Trainer
        GB, S = 256, 48
        lo, hi = D.shard_range(GB, rank, world)
        tr = Trainer(eng)
        g = tr.graph(hi - lo, S, S)
        gen = torch.Generator(device="cuda").manual_seed(7)
        xs = torch.rand((GB, S, S, 3), device="cuda", generator=gen)       # the same global batch on every rank
        ys = torch.rand((GB, 4 * S, 4 * S, 3), device="cuda", generator=gen)
        g.x_in.copy_(xs[lo:hi])
        g.y_true.copy_(ys[lo:hi])
        del xs, ys
        kt = max(3, min(steps, 10))

        def run_steps(fn):
            for _ in range(3):
                fn()
            D.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(kt):
                fn()
            e1.record()
            torch.cuda.synchronize()
            D.barrier()
            return D.max_over_ranks(e0.elapsed_time(e1) / kt)

        def step_no_comm():                  # the same step without any exchange (every rank on its own): the floor
            tr.forward_backward_device(g)
            tr.apply_gradients(summed_over=world)

        # the default step: world > 1 -> backward, then ONE fused reduce-scatter + Adam + all-gather kernel per rank over
        # NVLink peer memory (csrc/exchange.cu); the same step with ncclAllReduce + full-arena Adam is timed beside it
        ms = run_steps(lambda: tr.step_device(g))
        ms_nccl, p2p = None, tr.exchange is not None
        if p2p:
            ex, tr.exchange = tr.exchange, None
            ms_nccl = run_steps(lambda: tr.step_device(g))
            tr.exchange = ex
            assert not ex.timed_out(), "peer-memory exchange: a wait on a peer expired"
        ms_overlap = None
        if world > 1 and eng.sequencer == "python":   # the bucketed overlap needs the Python launch lists
            ms_overlap = run_steps(lambda: tr.step_device(g, overlap=True))
        ms_floor = run_steps(step_no_comm) if world > 1 else ms
        upd_ms = ms - ms_floor
        flops = tr.step_flops(g)
        line["train_step"] = {
            "workload": "training step (BASELINE configs[3]): 48x48 LR -> 192x192, global batch 256 split over the ranks, "
                        "forward + dgrad/wgrad bf16 on the tensor cores, gradient exchange + Keras-Adam (see comm), "
                        "weight repack", "global_batch": GB, "per_gpu_batch": hi - lo,
            "ms_per_step": round(ms, 3), "images_per_s": round(GB / ms * 1e3, 1), "steps": kt,
            "exchange": "p2p" if p2p else ("nccl" if world > 1 else "none"),
            "ms_per_step_nccl_allreduce_then_adam": None if ms_nccl is None else round(ms_nccl, 3),
            "ms_per_step_bucketed_overlapped_allreduce": None if ms_overlap is None else round(ms_overlap, 3),
            "sequencer": "sr_model_forward_backward + sr_model_apply_gradients (libsr100)" if eng.sequencer == "c"
                         else "python launch lists",
            "ms_per_step_without_exchange": round(ms_floor, 3),
            "exposed_exchange_ms": round(upd_ms, 3),
            "exposed_exchange_ms_nccl": None if ms_nccl is None else round(ms_nccl - ms_floor, 3),
            "exchange_bytes_per_rank": int(2 * (world - 1) / world * tr.grads.numel() * 4) if world > 1 else 0,
            "comm": tr.comm_description() if hasattr(tr, "comm_description") else "one all_reduce after backward",
            "algorithmic_tflop_per_step_per_gpu": round(flops / 1e12, 3),
            "tflops_per_gpu": round(flops / (ms * 1e-3) / 1e12, 1),
            "frac_of_sustained_peak": round(flops / (ms * 1e-3) / 1e12 / peaks["sustained"], 4),
            "loss": tr.last_loss(g) if hasattr(tr, "last_loss") else float(g.loss_sum.item()) / g.n_local,
            "scaling": "strong"}
        del tr, g
        torch.cuda.empty_cache()

    if rank != 0:
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1 and not args.profile:   # rank 0 at N = 1 only (~10-15 s of CPU work)
        threads = os.cpu_count() or 1
        v, sample, _ = cpu_reference_rate(threads)
        cpu = {"value": round(v, 5), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    out = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": round(t_res / steps * 1e3, 3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": eng.precision, "data": "synthetic",
        "config": headline_config(),
        "detail": {"images_this_rank": len(my_ids), "tiles_run_this_rank": tiles_run, "hr_stage_extents": hr_shapes,
                   "dead_work_elimination": "tiles that own no pixel of the final 4Hx4W image are not run (48 of 54 per "
                                            "image are); the HR stage runs on the 272x272 corner of each 384x384 patch "
                                            "that the stitch can see (receptive-field radius 7); output pixels "
                                            "bit-identical to the full tiling (tests/test_gpu_deadwork.py, "
                                            "tests/test_tile_plan.py), which is timed as value_full_tiles",
                   "residual_stream": "fp32 at LR and HR (tf32 operands)" if eng.tf32 else "fp32 at LR, bf16 at HR",
                   "sequencer": "sr_model_forward (libsr100 owns launch order, plans, CUDA graphs)" if eng.sequencer == "c"
                                else "python launch lists",
                   "stitch": "x255 / clip / uint8 / 8-px-crop stitch fused into the tail convs' epilogues" if fused_stitch
                             else "separate sr_patch_stitch pass per image"},
        "e2e": {"value": round(e2e, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "models.DifvdsrDouble.upscale_arrays (pinned host uint8 in / out, per rank)"},
        "value_full_tiles": None if value_full is None else round(value_full, 3),
        "gpu_launches": launches_total,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    out.update(line)
    print(json.dumps(out), flush=True)


def _shutdown():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        pass


if __name__ == "__main__":
    try:
        main()
    finally:
        _shutdown()
