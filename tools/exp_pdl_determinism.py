import json, os, sys
sys.path.insert(0, "/root/repo/image-enhance-keras_b200")
import torch
from sr100 import _lib as L
from sr100.engine import Engine, glorot_uniform_weights
from sr100.train import Trainer
lib = L.require_device()
w = glorot_uniform_weights(seed=1234)
res = []
for mode in (1, 0, 1, 0):
    lib.sr_set_pdl(mode)
    eng = Engine(w)
    tr = Trainer(eng)
    g = tr.graph(32, 48, 48)
    gen = torch.Generator(device="cuda").manual_seed(7)
    g.x_in.copy_(torch.rand(g.x_in.shape, device="cuda", generator=gen))
    g.y_true.copy_(torch.rand(g.y_true.shape, device="cuda", generator=gen))
    outs = []
    for i in range(4):
        tr.forward_backward_device(g)
        torch.cuda.synchronize()
        outs.append(tr.grads.clone())
    res.append(outs)
    del tr, g, eng
ref = res[0][0]
scale = float(ref.abs().max())
out = {}
for a in range(4):
    for i in range(4):
        d = (res[a][i] - ref).abs()
        out["m%d_r%d" % (a, i)] = (round(float(d.max()) / scale, 9), int((d > 0).sum()))
print(json.dumps(out))
