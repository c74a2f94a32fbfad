"""Host-side batching logic (CPU): the prefetched epoch orders equal the reference DataLoader's."""
import torch

from probaforms_b200.batching import PermutationPrefetcher, epoch_permutation, batch_bounds, shard_bounds


def test_prefetcher_reproduces_the_sequential_rng_consumption():
    n, epochs = 1000, 4
    torch.manual_seed(123)
    want = [epoch_permutation(n) for _ in range(epochs)]
    after_want = torch.rand(3)
    torch.manual_seed(123)
    pf = PermutationPrefetcher(n, epochs)
    got = [pf.next() for _ in range(epochs)]
    after_got = torch.rand(3)
    for a, b in zip(want, got):
        assert torch.equal(a, b)
    assert torch.equal(after_want, after_got)          # no extra draws from the global generator
    try:
        pf.next()
        assert False, "expected RuntimeError"
    except RuntimeError:
        pass


def test_prefetcher_matches_torch_dataloader_order():
    """The reference iterates DataLoader(TensorDataset, batch_size, shuffle=True) afresh every epoch (realnvp.py:237)."""
    from torch.utils.data import DataLoader, TensorDataset
    n, bs, epochs = 257, 32, 3
    data = torch.arange(n)
    torch.manual_seed(7)
    want = []
    for _ in range(epochs):
        want.append(torch.cat([b[0] for b in DataLoader(TensorDataset(data), batch_size=bs, shuffle=True)]))
    torch.manual_seed(7)
    pf = PermutationPrefetcher(n, epochs)
    for e in range(epochs):
        perm = pf.next()
        got = torch.cat([data[perm[b0:b0 + nb]] for b0, nb in batch_bounds(n, bs)])
        assert torch.equal(got, want[e])


def test_shards_tile_the_batch():
    for nb in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(10, nb, r, world) for r in range(world)]
            assert cuts[0][0] == 10 and cuts[-1][1] == 10 + nb
            assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))


def test_native_streaming_permutation_is_torch_randperm():
    """rnvp_perm_* (C ABI, host code) reproduces torch.randperm bit for bit, also when consumed in pieces."""
    from probaforms_b200 import _lib
    from probaforms_b200.batching import StreamingPermutation
    lib = _lib.load()
    for seed, n in [(0, 1), (1, 2), (12345, 1000), (2 ** 40 + 77, 65537), (2 ** 63 - 1, 300001)]:
        g = torch.Generator()
        g.manual_seed(seed)
        want = torch.randperm(n, generator=g)
        sp = StreamingPermutation(lib, seed, n, chunk=4099)
        first = sp.wait(min(n, 10))[: min(n, 10)].clone()          # a prefix is final before the rest
        assert torch.equal(first, want[: min(n, 10)])
        assert torch.equal(sp.full(), want)


def test_streaming_prefetcher_equals_sequential():
    from probaforms_b200 import _lib
    lib = _lib.load()
    n, epochs = 5000, 3
    torch.manual_seed(99)
    want = [epoch_permutation(n) for _ in range(epochs)]
    tail = torch.rand(2)
    torch.manual_seed(99)
    pf = PermutationPrefetcher(n, epochs, lib=lib)
    got = [pf.next_stream().full().clone() for _ in range(epochs)]
    assert all(torch.equal(a, b) for a, b in zip(want, got))
    assert torch.equal(tail, torch.rand(2))
