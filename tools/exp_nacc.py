"""Experiment: the headline workload (64 x 339x510, tiled) with two accumulators per weight stage and double-buffered
TMEM (nacc=2, the default) against four accumulators and no epilogue overlap (nacc=4: half the weight-stage fills),
inside a long power-capped run.  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bench import synth_image, B_H, B_W
    from sr100.engine import Engine, glorot_uniform_weights
    n_img = int(os.environ.get("N_IMG", "64"))
    imgs = [torch.from_numpy(synth_image(100 + i, B_H, B_W)).cuda() for i in range(n_img)]
    w = glorot_uniform_weights(seed=1234)
    out = {}
    for rep in range(2):
        for nacc in (2, 4):
            eng = Engine(w, nacc=nacc)
            for _ in range(2):
                eng.upscale_images_device(imgs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            k = 3
            for _ in range(k):
                eng.upscale_images_device(imgs)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / k
            out.setdefault("nacc%d_ms" % nacc, []).append(round(ms, 2))
            fl = eng.last_flops()
            out.setdefault("nacc%d_tflops" % nacc, []).append(round(fl / ms / 1e9, 1))
            eng.release()
            del eng
            torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
