"""world_size-2 ``gloo`` test of the data-parallel fit logic on CPU.

The product's host logic (probaforms_b200/batching.py: broadcast epoch permutation, batch and
shard bounds, one flat all-reduce of [gradients | loss]) is exercised with two processes; the
per-shard gradient sums come from the CPU oracle standing in for the CUDA kernel, exactly as the
kernel is called in ``FlowEngine.fit_step`` (scale = -1/B_global).  Result must equal the
single-process reference trajectory (semantics = reference with batch_size = B_global).
"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import realnvp_oracle as O
    from probaforms_b200.batching import epoch_permutation, batch_bounds, shard_bounds

    D, Cd, L, hidden, act = 5, 3, 4, (10,), "tanh"
    n, bs, lr = 101, 32, 0.01
    params = O.init_params(D, Cd, L, hidden, seed=0)            # replicated weights
    order = O.param_order(L, len(hidden))
    g = torch.Generator().manual_seed(1)
    X, C = torch.randn(n, D, generator=g), torch.randn(n, Cd, generator=g)
    torch.manual_seed(100 + rank)                               # ranks deliberately hold different RNG states
    st = O.AdamState(params, lr=lr)
    losses = []
    for _ in range(2):
        perm = epoch_permutation(n)                             # rank 0's sampler seed is broadcast
        for b0, nb in batch_bounds(n, bs):
            lo, hi = shard_bounds(b0, nb, rank, world)
            idx = perm[lo:hi]
            buf = torch.zeros(sum(params[k].numel() for k in order) + 1)
            if hi > lo:
                # what rnvp_backward accumulates for this shard: d/dtheta of (-1/nb) * sum_rows logp, and sum logp
                leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
                lp = O.flow_forward_rows(X[idx], C[idx], leaves, L, len(hidden), act)[2]
                grads = torch.autograd.grad((-1.0 / nb) * lp.sum(), [leaves[k] for k in order], allow_unused=True)
                flat = [torch.zeros_like(params[k]) if gr is None else gr for k, gr in zip(order, grads)]
                buf[:-1] = torch.cat([f.reshape(-1) for f in flat])
                buf[-1] = lp.sum().detach()
            dist.all_reduce(buf)                                # one bucket: gradients + loss
            off, gd = 0, {}
            for k in order:
                m = params[k].numel()
                gd[k] = buf[off:off + m].view_as(params[k])
                off += m
            O.adam_step(params, gd, st)
            losses.append(float(buf[-1]) * (-1.0 / nb))
    out[rank] = (np.array(losses), {k: v.numpy().copy() for k, v in params.items()}, perm.numpy().copy())
    dist.destroy_process_group()


def test_two_rank_fit_equals_single_process():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (l0, p0, perm0), (l1, p1, perm1) = out[0], out[1]
    assert np.array_equal(perm0, perm1)                         # same epoch order on both ranks
    assert np.array_equal(l0, l1)
    for k in p0:
        assert np.array_equal(p0[k], p1[k])                     # replicas stay bit-identical

    # single process, the reference's trajectory with batch_size = global batch and rank 0's RNG
    from oracle import realnvp_oracle as O
    torch.set_num_threads(1)
    D, Cd, L, hidden, act = 5, 3, 4, (10,), "tanh"
    params = O.init_params(D, Cd, L, hidden, seed=0)
    g = torch.Generator().manual_seed(1)
    X, C = torch.randn(101, D, generator=g), torch.randn(101, Cd, generator=g)
    torch.manual_seed(100)
    hist, _ = O.fit(X, C, params, L, len(hidden), act, batch_size=32, n_epochs=2, lr=0.01)
    ref = np.array([float(h) for h in hist])
    assert l0.shape == ref.shape
    assert np.allclose(l0, ref, rtol=1e-5, atol=1e-6)
    for k in params:
        assert np.allclose(p0[k], params[k].numpy(), rtol=1e-4, atol=2e-6), k


def test_shard_bounds_partition_every_batch():
    from probaforms_b200.batching import batch_bounds, shard_bounds
    for n, bs, world in [(1000, 32, 8), (101, 32, 2), (7, 16, 4), (65536 * 8, 65536, 8)]:
        seen = 0
        for b0, nb in batch_bounds(n, bs):
            edges = [shard_bounds(b0, nb, r, world) for r in range(world)]
            assert edges[0][0] == b0 and edges[-1][1] == b0 + nb
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            assert max(h - l for l, h in edges) - min(h - l for l, h in edges) <= 1
            seen += nb
        assert seen == n
