"""Dataset preparation (img_utils.transform_images): the per-image pixel pipeline on the device
(sr100.dataprep.transform_image_device) next to the same pipeline through the libraries the reference calls
(Pillow resize / SHARPEN, scipy.ndimage.gaussian_filter, numpy bytescale) on the host.  PNG encode/decode excluded
on both sides.

    python tools/bench_dataprep.py [--images 64] [--out gpurun_out/bench_dataprep.jsonl]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))
sys.path.insert(0, ROOT)


def host_pipeline(img, sf, true_upscale):
    """transform_images' body with PIL / scipy.ndimage doing the work (what the reference runs)."""
    import numpy as np
    from PIL import Image, ImageFilter
    from scipy.ndimage import gaussian_filter

    def bytescale(a):
        cmin, cmax = a.min(), a.max()
        cs = cmax - cmin
        return (((a - cmin) * (255.0 / (cs if cs else 1))).clip(0, 255) + 0.5).astype(np.uint8)

    im = np.asarray(Image.fromarray(img).resize((256, 256), resample=2))
    im = np.asarray(Image.fromarray(im).filter(ImageFilter.SHARPEN))
    hr = 16 * sf
    grid = [(x, y) for x in range(0, 256 - hr, 16) for y in range(0, 256 - hr, 16)]
    ys, xs = [], []
    for i in range(256):
        x, y = grid[i % len(grid)]
        ip = im[x:x + hr, y:y + hr].astype(np.float64)
        ys.append(bytescale(ip))
        op = Image.fromarray(bytescale(gaussian_filter(ip, sigma=0.5))).resize((16, 16), resample=3)
        if not true_upscale:
            op = op.resize((hr, hr), resample=3)
        xs.append(np.asarray(op))
    return np.stack(ys), np.stack(xs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_dataprep.jsonl"))
    a = ap.parse_args()
    import numpy as np
    import torch
    from sr100 import dataprep
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, size=(int(rng.integers(200, 500)), int(rng.integers(200, 500)), 3)).astype(np.uint8)
            for _ in range(a.images)]
    recs = []
    for sf, tu in ((2, False), (4, True)):
        dev = [torch.from_numpy(im).cuda() for im in imgs]
        for im in dev[:4]:
            dataprep.transform_image_device(im, sf, tu)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for im in dev:
            y, x = dataprep.transform_image_device(im, sf, tu)
        e1.record()
        torch.cuda.synchronize()
        ms_dev = e0.elapsed_time(e1) / len(dev)
        t0 = time.perf_counter()
        for im in imgs:                                  # host arrays in, host arrays out
            y, x = dataprep.transform_image_device(torch.from_numpy(im).cuda(), sf, tu)
            yh, xh = y.cpu().numpy(), x.cpu().numpy()
        ms_e2e = (time.perf_counter() - t0) * 1e3 / len(imgs)
        nh = min(8, len(imgs))
        t0 = time.perf_counter()
        for im in imgs[:nh]:
            wy, wx = host_pipeline(im, sf, tu)
        ms_host = (time.perf_counter() - t0) * 1e3 / nh
        same = bool(np.array_equal(wy, dataprep.transform_image_device(torch.from_numpy(imgs[nh - 1]).cuda(), sf, tu)[0].cpu().numpy()))
        rec = dict(workload="transform_images body, %d images of 200-500 px, scaling_factor %d, true_upscale %s" % (len(imgs), sf, tu),
                   device_ms_per_image=round(ms_dev, 3), e2e_ms_per_image_host_arrays=round(ms_e2e, 3),
                   host_pil_scipy_ms_per_image=round(ms_host, 2), host_sample=nh, launches_per_image=7 if not tu else 5,
                   speedup_e2e=round(ms_host / ms_e2e, 1), outputs_equal_host=same)
        recs.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        for r in recs:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
