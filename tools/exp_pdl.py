"""A/B of programmatic dependent launch (sr_set_pdl) on one GPU, interleaved so that box drift cancels: BASELINE
config 1 (one 128x128 patch, latency) and the training step at 32 patches (the per-GPU shape of the 8-GPU run) and
at 256.  Both variants are built first (their CUDA graphs are captured under their own setting), then timed
alternately.  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def timed(fn, n, warm):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    import torch
    from sr100 import _lib as L
    from sr100.engine import Engine, glorot_uniform_weights
    from sr100.train import Trainer
    lib = L.require_device()
    w = glorot_uniform_weights(seed=1234)
    x1 = torch.rand(1, 128, 128, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    rec, var, outs = {}, {}, {}
    for pdl in (1, 0):
        lib.sr_set_pdl(pdl)
        eng = Engine(w)
        for _ in range(5):
            out = eng.forward_device(x1)    # eager, capture, replay: the graph is built under this setting
        outs[pdl] = out.clone()             # (before any optimizer step changes the weights)
        tr = Trainer(eng)
        gs = {}
        for nb in (32, 256):
            g = tr.graph(nb, 48, 48)
            gen = torch.Generator(device="cuda").manual_seed(7)
            g.x_in.copy_(torch.rand(g.x_in.shape, device="cuda", generator=gen))
            g.y_true.copy_(torch.rand(g.y_true.shape, device="cuda", generator=gen))
            for _ in range(4):
                tr.step_device(g)
            gs[nb] = g
        var[pdl] = (eng, tr, gs)
    torch.cuda.synchronize()
    rec["config1_bit_equal"] = bool(torch.equal(outs[0], outs[1]))
    for rep in range(4):
        for pdl in (1, 0):
            eng, tr, gs = var[pdl]
            rec.setdefault("config1_ms_pdl%d" % pdl, []).append(round(timed(lambda: eng.forward_device(x1), 200, 20), 4))
            rec.setdefault("train_nb32_ms_pdl%d" % pdl, []).append(round(timed(lambda: tr.step_device(gs[32]), 20, 3), 3))
            if rep < 2:
                rec.setdefault("train_nb256_ms_pdl%d" % pdl, []).append(round(timed(lambda: tr.step_device(gs[256]), 4, 1), 2))
    lib.sr_set_pdl(1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
