"""BASELINE config 1 (one 128x128 patch -> 512x512 through model.predict): latency with the LR stage as one
persistent chain launch (default) and as 60 per-layer launches (SR100_CHAIN_LR=0).  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def main():
    import torch
    from sr100.engine import Engine, glorot_uniform_weights
    w = glorot_uniform_weights(seed=1234)
    x = torch.rand(1, 128, 128, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    rec = {}
    for rep in range(2):
        for chain in ("1", "0"):
            os.environ["SR100_CHAIN_LR"] = chain
            eng = Engine(w)
            for _ in range(20):
                eng.forward_device(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                eng.forward_device(x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 200
            rec.setdefault("chain%s_ms" % chain, []).append(round(ms, 4))
            rec.setdefault("chain%s_tflops" % chain, []).append(round(eng.last_flops() / ms / 1e9, 1))
            del eng
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
