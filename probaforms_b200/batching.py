"""Host-side batching logic of ``RealNVP.fit`` (reference realnvp.py:229-237), device-free.

Kept separate from the CUDA plumbing so that the data-parallel arithmetic can be tested with the
``gloo`` backend on CPU (tests/test_dist_gloo.py).
"""
import torch


def epoch_permutation(n, group=None, device=None):
    """Row order of one epoch, consuming the global torch RNG exactly as the reference's fresh
    ``DataLoader(dataset, batch_size, shuffle=True)`` does (realnvp.py:237): one int64 draw for the
    loader's base seed (torch/utils/data/dataloader.py ``_BaseDataLoaderIter.__init__``), one for
    the ``RandomSampler`` seed (sampler.py ``RandomSampler.__iter__``), then ``randperm(n)`` from a
    generator seeded with the latter.  With a process group the sampler seed of rank 0 is
    broadcast so every rank walks the same order (ranks may hold different RNG states)."""
    torch.empty((), dtype=torch.int64).random_()
    seed = torch.empty((), dtype=torch.int64).random_()
    if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                             and torch.distributed.get_world_size() > 1):
        s = seed.reshape(1).to(device) if device is not None else seed.reshape(1)
        torch.distributed.broadcast(s, src=0, group=group)
        seed = s.cpu().reshape(())
    g = torch.Generator()
    g.manual_seed(int(seed.item()))
    return torch.randperm(n, generator=g)


def batch_bounds(n, batch_size):
    """[(b0, nb)] of consecutive batches; the last partial batch is kept (drop_last=False)."""
    return [(b0, min(batch_size, n - b0)) for b0 in range(0, n, batch_size)]


def shard_bounds(b0, nb, rank, world):
    """This rank's contiguous slice [lo, hi) of the global batch [b0, b0+nb): near-equal shards
    whose union is exactly the batch, so the all-reduced gradient sum equals the single-process one."""
    return b0 + (nb * rank) // world, b0 + (nb * (rank + 1)) // world
