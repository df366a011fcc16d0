"""Experiment: training step (forward + backward through sr_model_forward_backward, Adam) with the independent
launches paired on two streams (sr_model_config.overlap_train) against the single-stream sequence, at the per-GPU
minibatches of the 1-, 2- and 8-GPU runs of BASELINE config 4 (256 / 128 / 32 patches of 48x48).  JSON lines."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def main():
    import torch
    from sr100.engine import Engine, glorot_uniform_weights
    from sr100.train import Trainer
    w = glorot_uniform_weights(seed=1234)
    for nb in (32, 128, 256):
        rec = {"per_gpu_batch": nb}
        for rep in range(2):
            for ov in ("0", "1"):
                os.environ["SR100_OVERLAP_TRAIN"] = ov
                eng = Engine(w)
                tr = Trainer(eng)
                g = tr.graph(nb, 48, 48)
                gen = torch.Generator(device="cuda").manual_seed(7)
                g.x_in.copy_(torch.rand(g.x_in.shape, device="cuda", generator=gen))
                g.y_true.copy_(torch.rand(g.y_true.shape, device="cuda", generator=gen))
                for _ in range(3):
                    tr.step_device(g)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k = 6
                e0.record()
                for _ in range(k):
                    tr.step_device(g)
                e1.record()
                torch.cuda.synchronize()
                rec.setdefault("overlap%s_ms" % ov, []).append(round(e0.elapsed_time(e1) / k, 3))
                rec["loss_%s" % ov] = tr.last_loss(g)
                del tr, g, eng
                torch.cuda.empty_cache()
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
