"""The training exchange alone (no forward / backward): the fused reduce-scatter + Adam + all-gather kernel over NVLink
peer memory (csrc/exchange.cu) against what it replaces, ncclAllReduce(sum) of the gradient arena + the full-arena
Adam launch, on the real arena size (21,838,211 floats).  Back-to-back steps between CUDA events, max over ranks.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_exchange.py [--steps 50]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-enhance-keras_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--n", type=int, default=21838211)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from sr100 import _lib as L, dist as D, peer
    rank, local_rank, world = D.init_process_group()
    torch.cuda.set_device(local_rank if world > 1 else 0)
    lib = L.require_device()
    n = a.n
    gen = torch.Generator(device="cuda").manual_seed(3)
    p0 = torch.randn(n, device="cuda", generator=gen)
    g0 = torch.randn(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(10 + rank))
    hp = (1e-4, 0.9, 0.999, 1e-7)

    def timed(fn):
        for _ in range(5):
            fn()
        D.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        D.barrier()
        return D.max_over_ranks(e0.elapsed_time(e1) / a.steps)

    # --- fused kernel over peer memory
    params, grads = p0.clone(), g0.clone()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    t = [0]
    if world > 1:
        ex = peer.connect(lib, grads, params)
    else:
        sig = torch.zeros(lib.sr_exchange_signal_bytes() // 4, dtype=torch.int32, device="cuda")
        ex = peer.Exchange(lib, 0, 1, n, [grads.data_ptr()], [params.data_ptr()], [sig.data_ptr()])

    def step_p2p():
        t[0] += 1
        ex.adam_step(m, v, t[0], *hp, 1.0 / world)

    ms_p2p = timed(step_p2p)
    assert not ex.timed_out()
    # --- all-reduce + full-arena Adam
    params2, grads2 = p0.clone(), g0.clone()
    m2, v2 = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    t2 = [0]

    def step_nccl():
        t2[0] += 1
        if world > 1:
            dist.all_reduce(grads2)
        L.check(lib.sr_adam_step(L.ptr(params2), L.ptr(grads2), L.ptr(m2), L.ptr(v2), n, *hp, t2[0], 1.0 / world,
                                 L.stream_ptr()))

    ms_nccl = timed(step_nccl)

    def step_allreduce_only():
        if world > 1:
            dist.all_reduce(grads2)

    ms_ar = timed(step_allreduce_only) if world > 1 else 0.0
    # one step of each from the same state: same parameters up to the fp32 summation order of the all-reduce
    params.copy_(p0); params2.copy_(p0); grads.copy_(g0); grads2.copy_(g0)
    for b in (m, v, m2, v2):
        b.zero_()
    torch.cuda.synchronize()
    D.barrier()
    ex.adam_step(m, v, 1, *hp, 1.0 / world)
    if world > 1:
        dist.all_reduce(grads2)
    L.check(lib.sr_adam_step(L.ptr(params2), L.ptr(grads2), L.ptr(m2), L.ptr(v2), n, *hp, 1, 1.0 / world, L.stream_ptr()))
    torch.cuda.synchronize()
    diff = float((params - params2).abs().max())
    allp = [torch.empty(1024, device="cuda") for _ in range(world)]
    if world > 1:
        dist.all_gather(allp, params[:1024].contiguous())
    same = all(torch.equal(allp[0], q) for q in allp) if world > 1 else True
    if rank == 0:
        wire = 2.0 * (world - 1) / world * n * 4
        print(json.dumps(dict(metric="training_exchange", n_gpus=world, floats=n, steps=a.steps,
                              ms_fused_p2p=round(ms_p2p, 4), ms_nccl_allreduce_plus_adam=round(ms_nccl, 4),
                              ms_nccl_allreduce_only=round(ms_ar, 4),
                              p2p_wire_bytes_per_rank=int(wire),
                              p2p_link_GBps_per_direction=round(wire / 2 / (ms_p2p * 1e-3) / 1e9, 1) if world > 1 else None,
                              max_abs_param_diff_vs_nccl=diff, replicas_bit_equal=bool(same))), flush=True)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
