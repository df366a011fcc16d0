"""Write the flat fp32 parameter arena of the graph-level C ABI (sr_model_create: kernel HWIO then bias per layer, in
sr_model_layer order) from a Keras weight file (.h5 / .npz) or from a seeded glorot_uniform initialisation -- the input
of examples/sr_predict.c.  Host only (no GPU needed).

    python tools/export_arena.py --out params.f32 [--weights file.h5|file.npz] [--seed 1234]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def arena_from_dict(weights):
    from sr100.engine import layer_specs
    parts = []
    for name, k, cin, cout in layer_specs():
        w, b = weights[name]
        assert w.shape == (k, k, cin, cout) and b.shape == (cout,), name
        parts += [np.ascontiguousarray(w, dtype=np.float32).ravel(), np.ascontiguousarray(b, dtype=np.float32)]
    return np.concatenate(parts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--weights")
    ap.add_argument("--seed", type=int, default=1234)
    a = ap.parse_args()
    from sr100.engine import glorot_uniform_weights, layer_specs
    if a.weights:
        names = [s[0] for s in layer_specs()]
        if a.weights.endswith(".npz"):
            z = np.load(a.weights)
            w = {n: (z[n + "/kernel:0"], z[n + "/bias:0"]) for n in names}
        else:
            from sr100 import h5lite
            w = h5lite.load_keras_weights(a.weights, names)
    else:
        w = glorot_uniform_weights(seed=a.seed)
    arena = arena_from_dict(w)
    arena.tofile(a.out)
    print("%s: %d floats" % (a.out, arena.size))


if __name__ == "__main__":
    main()
