"""-m gpu: the fused reduce-scatter + Adam + all-gather kernel of the training exchange (csrc/exchange.cu,
include/sr100.h sr_exchange_*; the optimizer is compile(Adam(1e-4, 0.9)), models.py:1212-1213).

Checked against what it replaces -- the sum of the ranks' gradient arenas (rank order) followed by sr_adam_step over the
whole arena -- bit for bit: (1) W logical ranks inside one process (W kernels on W streams over W sets of arenas, small
grids so that they are resident together), (2) two PROCESSES on the one GPU of the test box with the arenas mapped
through CUDA IPC (the path bench.py --gpus N runs over NVLink)."""
import ctypes as C
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HP = dict(lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7)


def _reference(lib, p0, grads_per_step, world):
    """sum in rank order -> sr_adam_step on the whole arena, step after step."""
    from sr100 import _lib as L
    p = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t, gs in enumerate(grads_per_step, 1):
        g = gs[0].clone()
        for q in gs[1:]:
            g += q
        L.check(lib.sr_adam_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), p.numel(), HP["lr"], HP["beta_1"],
                                 HP["beta_2"], HP["epsilon"], t, 1.0 / world, L.stream_ptr()))
    torch.cuda.synchronize()
    return p, m, v


@pytest.mark.parametrize("world,n", [(1, 1003), (2, 4099), (3, 70001), (8, 1000003), (8, 37)])
def test_logical_ranks_equal_allreduce_then_adam(lib, world, n):
    from sr100 import peer
    gen = torch.Generator(device="cuda").manual_seed(world * 1000 + n)
    p0 = torch.randn(n, device="cuda", generator=gen)
    params = [p0.clone() for _ in range(world)]
    grads = [torch.zeros(n, device="cuda") for _ in range(world)]
    sig = [torch.zeros(lib.sr_exchange_signal_bytes() // 4, dtype=torch.int32, device="cuda") for _ in range(world)]
    ms = [torch.zeros(n, device="cuda") for _ in range(world)]
    vs = [torch.zeros(n, device="cuda") for _ in range(world)]
    exs = [peer.Exchange(lib, r, world, n, [g.data_ptr() for g in grads], [p.data_ptr() for p in params],
                         [s.data_ptr() for s in sig]) for r in range(world)]
    shards = [(e.lo, e.hi) for e in exs]
    assert shards[0][0] == 0 and shards[-1][1] == n and all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
    streams = [torch.cuda.Stream() for _ in range(world)]
    steps = []
    for t in range(1, 4):
        gs = [torch.randn(n, device="cuda", generator=gen) * (0.5 + r) for r in range(world)]
        steps.append(gs)
        for r in range(world):
            grads[r].copy_(gs[r])
        torch.cuda.synchronize()
        for r in range(world):               # the W kernels wait for each other: small grids, one stream each
            exs[r].set_timeout_ms(5000)
            with torch.cuda.stream(streams[r]):
                exs[r].adam_step(ms[r], vs[r], t, HP["lr"], HP["beta_1"], HP["beta_2"], HP["epsilon"], 1.0 / world,
                                 max_blocks=4)
        torch.cuda.synchronize()
        assert not any(e.timed_out() for e in exs)
    p_ref, m_ref, v_ref = _reference(lib, p0, steps, world)
    for r in range(world):
        assert torch.equal(params[r], p_ref), "rank %d parameters differ from all-reduce + Adam" % r
        lo, hi = shards[r]
        assert torch.equal(ms[r][lo:hi], m_ref[lo:hi]) and torch.equal(vs[r][lo:hi], v_ref[lo:hi])
        if lo > 0:                           # the optimizer state outside the shard is never touched
            assert float(ms[r][:lo].abs().max()) == 0.0
        if hi < n:
            assert float(vs[r][hi:].abs().max()) == 0.0


def test_missing_peer_times_out_instead_of_hanging(lib):
    """Rank 0 of a world of 2 whose peer never launches: the bounded wait expires and the status says so."""
    from sr100 import peer
    n = 1024
    bufs = [torch.zeros(n, device="cuda") for _ in range(6)]
    sig = [torch.zeros(lib.sr_exchange_signal_bytes() // 4, dtype=torch.int32, device="cuda") for _ in range(2)]
    ex = peer.Exchange(lib, 0, 2, n, [bufs[0].data_ptr(), bufs[1].data_ptr()], [bufs[2].data_ptr(), bufs[3].data_ptr()],
                       [s.data_ptr() for s in sig])
    ex.set_timeout_ms(200)
    ex.adam_step(bufs[4], bufs[5], 1, 1e-3, 0.9, 0.999, 1e-7, 0.5, max_blocks=2)
    torch.cuda.synchronize()
    assert ex.timed_out()


IPC_WORKER = r"""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "image-enhance-keras_b200"))
torch.cuda.set_device(0)                      # both processes share the test box's one GPU
from sr100 import _lib as L, peer
lib = L.require_device()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group(backend="gloo", rank=rank, world_size=world)
n = %(n)d
gen = torch.Generator(device="cuda").manual_seed(5)
p0 = torch.randn(n, device="cuda", generator=gen)
pad = torch.empty(12345 + 1000 * rank, device="cuda")        # different offsets inside the allocator's blocks
params = p0.clone()
grads = torch.zeros(n, device="cuda")
m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
ex = peer.connect(lib, grads, params)
ex.set_timeout_ms(20000)
steps = []
for t in range(1, 4):
    gs = [torch.randn(n, device="cuda", generator=gen) * (0.5 + r) for r in range(world)]   # same stream on both ranks
    steps.append(gs)
    grads.copy_(gs[rank])
    ex.adam_step(m, v, t, 1e-3, 0.9, 0.999, 1e-7, 1.0 / world)
    torch.cuda.synchronize()
assert not ex.timed_out()
p = p0.clone(); mr = torch.zeros(n, device="cuda"); vr = torch.zeros(n, device="cuda")
for t, gs in enumerate(steps, 1):
    g = gs[0].clone()
    for q in gs[1:]:
        g += q
    L.check(lib.sr_adam_step(L.ptr(p), L.ptr(g), L.ptr(mr), L.ptr(vr), n, 1e-3, 0.9, 0.999, 1e-7, t, 1.0 / world, L.stream_ptr()))
torch.cuda.synchronize()
ok = bool(torch.equal(p, params)) and bool(torch.equal(m[ex.lo:ex.hi], mr[ex.lo:ex.hi]))
print(json.dumps(dict(rank=rank, ok=ok, lo=ex.lo, hi=ex.hi)))
dist.barrier()
dist.destroy_process_group()
"""


def test_two_processes_over_cuda_ipc(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = IPC_WORKER % dict(root=root, n=300007)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", script], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    recs = [json.loads([l for l in o[0].splitlines() if l.startswith("{")][-1]) for o in outs]
    assert all(r["ok"] for r in recs), recs
    recs.sort(key=lambda r: r["rank"])
    assert recs[0]["lo"] == 0 and recs[0]["hi"] == recs[1]["lo"] and recs[1]["hi"] == 300007


SHARD_WORKER = r"""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "image-enhance-keras_b200"))
torch.cuda.set_device(0)
from sr100 import dist as D
from sr100.engine import Engine, glorot_uniform_weights
rank, local_rank, world = D.init_process_group(backend="gloo")
eng = Engine(glorot_uniform_weights(seed=3))
rng = np.random.default_rng(4)
outs = []
for (h, w) in ((150, 330), (100, 200), (150, 330)):        # a smaller image after a larger one reuses the mapped canvas
    img = torch.from_numpy(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)).cuda()
    got = eng.upscale_image_sharded(img)
    torch.cuda.synchronize()
    if rank == 0:
        want = eng.upscale_image_sharded(img, world=1, rank=0)
        outs.append(bool(torch.equal(got, want)) and tuple(got.shape) == (4 * h, 4 * w, 3))
    else:
        assert got is None
    dist.barrier()
if rank == 0:
    st = eng._peer_canvas_state
    print(json.dumps(dict(equal=outs, p2p=st is not None and st != "unavailable", timed_out=st[3].timed_out())))
dist.barrier()
dist.destroy_process_group()
"""


def test_tile_sharded_image_written_into_rank0_over_peer_memory():
    """BASELINE config 5 with the peer-memory path on: two processes (ranks) on the test box's one GPU, rank 1's tail
    convs store their owned pixels through a CUDA-IPC mapping of rank 0's image; bit-identical to one rank."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = SHARD_WORKER % dict(root=root)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK="0", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), SR100_SHARD_GATHER="p2p")
        procs.append(subprocess.Popen([sys.executable, "-c", script], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    rec = json.loads([l for l in outs[0][0].splitlines() if l.startswith("{")][-1])
    assert rec["p2p"] and not rec["timed_out"] and rec["equal"] == [True, True, True], rec
