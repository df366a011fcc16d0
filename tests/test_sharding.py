"""CPU tests of the multi-GPU host logic: tile / image sharding and a world_size-2 gloo run
that checks shard-union == unsharded order and the max-over-ranks timing reduction used by bench.py."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_contiguous_shards_cover_everything():
    from sr100.dist import shard_range, shard_round_robin
    for n in (0, 1, 7, 54, 558, 3456):
        for world in (1, 2, 4, 8):
            got = []
            sizes = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                got.extend(range(lo, hi))
                sizes.append(hi - lo)
            assert got == list(range(n))
            assert max(sizes) - min(sizes) <= 1
            rr = sorted(i for r in range(world) for i in shard_round_robin(n, r, world))
            assert rr == list(range(n))
    # config 5: 558 tiles over 8 ranks = 70 x 6 + 69 x 2
    assert [shard_range(558, r, 8)[1] - shard_range(558, r, 8)[0] for r in range(8)] == [70] * 6 + [69] * 2


WORKER = r'''
import os, sys, json
sys.path.insert(0, os.path.join(%(root)r, "image-enhance-keras_b200"))
import torch, torch.distributed as dist
from sr100.dist import init_process_group, shard_range, max_over_ranks, gather_objects
init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(11, rank, world)
mine = [i * i for i in range(lo, hi)]
allv = gather_objects(mine)
t = max_over_ranks(float(rank + 1) * 1.5)
if rank == 0:
    flat = [v for part in allv for v in part]
    print(json.dumps(dict(flat=flat, tmax=t, world=world)))
dist.barrier()
dist.destroy_process_group()
'''


def test_world_size_2_gloo():
    import json
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = WORKER % dict(root=ROOT)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", script], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    rec = json.loads([l for l in outs[0][0].splitlines() if l.startswith("{")][-1])
    assert rec["flat"] == [i * i for i in range(11)]
    assert rec["tmax"] == 3.0 and rec["world"] == 2


DP_WORKER = r"""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "image-enhance-keras_b200"))
from sr100 import dist as D
from oracle import model as om
torch.set_num_threads(2)
rank, local_rank, world = D.init_process_group(backend="gloo")
weights = om.init_weights(3, bias_scale=0.01)
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.random((4, 6, 6, 3)).astype(np.float32))
y = torch.from_numpy(rng.random((4, 24, 24, 3)).astype(np.float32))
lo, hi = D.shard_range(4, rank, world)
m = om.DifvdsrDoubleOracle(weights)
params = list(m.parameters())
loss = om.mse_loss(m(x[lo:hi]), y[lo:hi])                 # per-rank mean over the local minibatch
flat = torch.cat([g.reshape(-1) for g in torch.autograd.grad(loss, params)])
w = D.all_reduce_sum_(flat)                               # the training step's exchange
flat /= w                                                 # (folded into sr_adam_step(grad_scale) on the GPU)
if rank == 0:
    full = om.mse_loss(m(x), y)
    ref = torch.cat([g.reshape(-1) for g in torch.autograd.grad(full, params)])
    rel = float((flat - ref).norm() / ref.norm())
    print(json.dumps(dict(rel=rel, world=w, n=int(flat.numel()))))
dist.barrier()
dist.destroy_process_group()
"""


def test_data_parallel_gradient_mean_world_2_gloo():
    """The exchange step of the training path on CPU (gloo, world 2): all-reduce(sum)/world of the per-rank
    mean-loss gradients == gradient of the mean loss over the global minibatch (oracle graph, autograd)."""
    import json
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = DP_WORKER % dict(root=ROOT)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", script], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    rec = json.loads([l for l in outs[0][0].splitlines() if l.startswith("{")][-1])
    assert rec["world"] == 2 and rec["n"] == 21838211
    assert rec["rel"] < 1e-5
