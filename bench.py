"""Benchmark of the x4 SR hot path (BASELINE.json metric: x4 output megapixels/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload = "set5_x4_tiled", BASELINE.json configs[1]): five synthetic RGB images of the Set5
shapes (512x512, 288x288, 256x256, 280x280, 344x228) -> x4 through the reference's tiling (96/64 patches,
81+25+25+25+30 = 186 tiles, 8-px crop stitch) and the 86-conv DifvdsrDouble stack with random-init weights,
then Y-PSNR / Y-SSIM / RGB-SSIM scoring of every output.  One step = one pass over the five images
(9.079 output MP).  N > 1: one process per GPU (torchrun), every rank runs the same workload on its own
images, no data-path collective (weak scaling); value = total MP / max-over-ranks device time.

value: inputs resident in HBM when the timed region starts.  e2e: the same step through the public API
(models.DifvdsrDouble.upscale_arrays + scorpath.score_pair) with host buffers: H2D of the uint8 inputs and
D2H of the uint8 outputs and the scores inside the timed region.
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
METRIC = "x4_output_megapixels_per_sec"
UNIT = "MP/s"


def synth_images(seed):
    import numpy as np
    from scipy.ndimage import uniform_filter
    rng = np.random.default_rng(seed)
    out = []
    for h, w in SET5_SHAPES:
        img = rng.integers(0, 256, size=(h + 4, w + 4, 3)).astype(np.float32)
        out.append(uniform_filter(img, size=(5, 5, 1))[2:-2, 2:-2].astype(np.uint8))   # 5x5 box blur: not white noise
    return out


def output_megapixels():
    return sum(16 * h * w for h, w in SET5_SHAPES) / 1e6


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
    """Oracle (CPU restatement of the reference graph; Keras/TF are not installable offline) on a bounded
    sample: the first n 96x96 tiles of the 256x256 'butterfly' image (25 tiles -> 1.049 output MP)."""
    import numpy as np
    import torch
    from oracle import model as om
    from oracle import tiling as ot
    torch.set_num_threads(threads)
    img = synth_images(0)[2]
    canvas = ot.make_canvas(img, 96, 64)
    patches, counts = ot.extract_patches_step(canvas, (96, 96), 64)
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
    mp = (n / float(patches.shape[0])) * (16 * 256 * 256 / 1e6)
    return mp / dt, "first %d of 25 96x96 tiles of the 256x256 image, torch CPU fp32 oracle, %.1f s" % (n, dt), dt


def run_reference(args):
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
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": output_megapixels() / value * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "set5_x4_tiled", "images": SET5_SHAPES, "tiles": 186, "patch": 96, "step": 64,
                   "output_mp_per_step_per_gpu": output_megapixels(), "weights": "glorot_uniform random init",
                   "sample": "each step times a bounded sample of the workload's tiles on the host cores and scales to "
                             "the full step (all 186 tiles of the literal reference tiling)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample + " per step; Keras/TensorFlow cannot be installed offline, so this is the "
                                            "CPU restatement of the reference graph (oracle/model.py)"},
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from sr100 import dist as D
    from sr100 import _lib as L
    from sr100 import ops
    rank, local_rank, world = D.init_process_group()
    torch.cuda.set_device(local_rank if world > 1 else 0)
    import models
    import scorpath

    def pinned(a):
        """Host arrays of the end-to-end path live in page-locked memory (what a serving loop would reuse)."""
        t = torch.empty(a.shape, dtype=torch.uint8, pin_memory=True)
        t.numpy()[...] = a
        return t.numpy()

    images = [pinned(im) for im in synth_images(100 + rank)]
    m = models.DifvdsrDouble(1)
    model = m.create_model(96, 96)                      # random-init glorot weights (no trained weights offline)
    eng = model.engine
    from sr100.engine import glorot_uniform_weights
    eng.set_weights_dict(glorot_uniform_weights(seed=1234))   # the same replica on every rank
    dev_imgs = [torch.from_numpy(im).cuda() for im in images]
    rng = np.random.default_rng(7 + rank)
    gts = [pinned(rng.integers(0, 256, size=(4 * h, 4 * w, 3)).astype(np.uint8)) for h, w in SET5_SHAPES]
    dev_gts = [torch.from_numpy(g).cuda() for g in gts]
    score_buf = torch.zeros(5, 64, dtype=torch.uint8, device="cuda")

    def step_resident(full=False):
        canv = eng.upscale_images_device(dev_imgs, full_canvas=full)
        score_buf.zero_()
        for i, (c, g) in enumerate(zip(canv, dev_gts)):
            h, w = g.shape[0], g.shape[1]
            sr = c[:h, :w].contiguous() if full else c
            L.check(eng.lib.sr_score_pair_u8(L.ptr(sr), L.ptr(g), h, w, 10, L.ptr(score_buf[i]), L.stream_ptr()))
        return canv

    def step_e2e():
        outs = m.upscale_arrays(images)
        return [scorpath.score_pair(g, o, 10) for g, o in zip(gts, outs)]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        D.barrier()
        return D.max_over_ranks(e0.elapsed_time(e1) / 1e3)

    sampler = ClockSampler(local_rank if world > 1 else 0)
    if rank == 0:
        sampler.start()
    t_res = timed(step_resident, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    # end to end: wall clock brackets the host work too (events alone would miss the D2H waits)
    for _ in range(2):
        step_e2e()
    D.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    t_e2e = D.max_over_ranks(time.perf_counter() - t0)

    mp_step = output_megapixels()
    value = world * mp_step * args.steps / t_res
    e2e = world * mp_step * args.steps / t_e2e

    # the reference's literal tile set (all 186 tiles, full 384x384 HR stage) for comparison: same output pixels
    t_full = timed(lambda: step_resident(True), max(2, args.steps // 2), 3)
    value_full = world * mp_step * max(2, args.steps // 2) / t_full

    # ---- roofline of the dominant kernel (conv_tc_pair_kernel): per-launch CUDA events over one more step
    from sr100.engine import _Plan
    step_resident()
    torch.cuda.synchronize()
    stages = list(eng.last_stages)
    tiles_run = sum(st_.NB for st_ in stages if hasattr(st_, "x_in"))
    hr_shapes = sorted({(st_.eh, st_.ew) for st_ in stages if hasattr(st_, "eh")})
    st = L.stream_ptr()
    evs = []
    for stg in stages:
        for step in stg.steps:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step(st)
            b.record()
            evs.append((a, b, isinstance(getattr(step, "__self__", None), _Plan)))
    torch.cuda.synchronize()
    conv_ms = sum(a.elapsed_time(b) for a, b, is_conv in evs if is_conv)
    n_conv = sum(1 for _, _, is_conv in evs if is_conv)
    total_ms = sum(a.elapsed_time(b) for a, b, _ in evs)
    conv_flops = eng.last_flops()
    peaks = load_peaks()
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12
    if eng.tf32:     # SR100_PRECISION=tf32: kind::tf32 MMAs run at half the bf16 rate; no measured tf32 peak exists
        peaks = {"sustained": peaks["sustained"] / 2, "burst": peaks["burst"] / 2,
                 "src": peaks["src"] + " (halved: tf32 = 0.5x bf16 tensor rate)"}
    launches_per_step = len(evs) + 5 + 5 + 5          # libsr100 kernels only: + gathers, stitches, scores
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
                "peak": peaks["sustained"], "unit": "TFLOP/s", "frac": round(achieved / peaks["sustained"], 4),
                "peak_source": peaks["src"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                "frac_of_burst_peak": round(achieved / peaks["burst"], 4), "traffic": traffic,
                "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
                "launches": n_conv, "avg_launch_ms": round(conv_ms / n_conv, 4),
                "share_of_forward": round(conv_ms / total_ms, 4),
                "executed_flops_per_step": conv_flops,
                "reference_tiling_flops_per_step": 189595215986688.0}

    if rank != 0:
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only (bounded sample, ~10-15 s of CPU work)
        threads = os.cpu_count() or 1
        v, sample, _ = cpu_reference_rate(threads)
        cpu = {"value": round(v, 5), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    # upscale_arrays: uint8 images in, cropped uint8 x4 images out (pinned); scorpath.score_pair: GT + SR in, scores out
    h2d = sum(im.nbytes for im in images) + sum(g_.nbytes for g_ in gts) + sum(16 * im.nbytes for im in images)
    d2h = sum(16 * im.nbytes for im in images) + 5 * 56
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(t_res / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": eng.precision, "data": "synthetic",
        "config": {"workload": "set5_x4_tiled", "images": SET5_SHAPES, "tiles": 186, "patch": 96, "step": 64,
                   "tiles_run": tiles_run, "hr_stage_extents": hr_shapes,
                   "dead_work_elimination": "tiles that own no pixel of the final 4Hx4W image are not run; the HR "
                                            "stage runs on the 272x272 corner of each 384x384 patch that the stitch "
                                            "can see (receptive-field radius 7); output pixels bit-identical to the "
                                            "full tiling (tests/test_gpu_deadwork.py, tests/test_tile_plan.py), which is timed as value_full_tiles",
                   "output_mp_per_step_per_gpu": mp_step, "weights": "glorot_uniform random init",
                   "residual_stream": "fp32 at LR and HR (tf32 operands)" if eng.tf32 else "fp32 at LR, bf16 at HR",
                   "l2": "every conv launch streams 0.36-2.9 GB of activations (> 126 MB L2); no flush needed"},
        "e2e": {"value": round(e2e, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "value_full_tiles": round(value_full, 3),
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


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
