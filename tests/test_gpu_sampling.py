"""In-kernel prior draws (rnvp_sample), sharded sampling / log-density and the ingestion paths (``-m gpu``)."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import realnvp_oracle as O

pytestmark = pytest.mark.gpu

SHAPES = [(2, 1, 8, (10,)), (32, 8, 4, (64,)), (64, 16, 3, (32,)), (7, 0, 5, (16, 12)), (5, 3, 4, (10,)), (24, 8, 3, (64,)),
          (9, 2, 3, (32,)), (128, 32, 2, (128,))]


def _flow(shape, seed, dev):
    from probaforms_b200.models import RealNVPLayer, NormalizingFlow
    D, Cd, L, hidden = shape
    params = O.init_params(D, Cd, L, hidden, seed=seed)
    nf = NormalizingFlow([RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, hidden, "tanh") for i in range(L)], prior=None)
    nf.load_state_dict(params)
    return nf.to(dev), params


@pytest.mark.parametrize("shape", SHAPES)
def test_in_kernel_noise_matches_the_philox_oracle(shape):
    """sample(C, seed) == oracle inverse pass on oracle.philox_normal(seed): all three kernel families, ragged N,
    and a row offset beyond 2^32 (the counter is the 64-bit global row index)."""
    dev = torch.device("cuda:0")
    D, Cd, L, hidden = shape
    nf, params = _flow(shape, 3, dev)
    for N, off in ((1000, 0), (333, 2 ** 33 + 5)):
        g = torch.Generator().manual_seed(N)
        Cn = torch.randn(N, Cd, generator=g) if Cd else None
        got = nf.sample(Cn.to(dev) if Cd else N, seed=1234567890123, row_offset=off).cpu()
        eps = O.philox_normal(1234567890123, off, N, D)
        want = O.flow_sample_from_noise(eps, Cn, params, L, len(hidden), "tanh")
        assert float((got - want).abs().max()) < 1e-5 * float(want.abs().max())
        # and against the same library's parity mode on the oracle's noise: the noise itself agrees to fp32 rounding
        same = nf.sample_from_noise(eps.to(dev), Cn.to(dev) if Cd else None).cpu()
        assert float((got - same).abs().max()) < 2e-6 * float(want.abs().max())


def test_sampling_is_independent_of_the_sharding():
    """Row blocks sampled separately (as different GPUs would) are exactly the rows of the single request."""
    dev = torch.device("cuda:0")
    nf, _ = _flow((32, 8, 4, (64,)), 5, dev)
    N = 5000
    Cn = torch.randn(N, 8, device=dev)
    whole = nf.sample(Cn, seed=99)
    parts = [nf.sample(Cn[a:b].contiguous(), seed=99, row_offset=a) for a, b in ((0, 1234), (1234, 1235), (1235, N))]
    assert torch.equal(torch.cat(parts), whole)
    torch.manual_seed(4)
    a = nf.sample(Cn)
    torch.manual_seed(4)
    assert torch.equal(nf.sample(Cn), a)                      # default seed comes from torch's global generator
    assert not torch.equal(nf.sample(Cn), a)


def test_model_sample_devices_and_log_prob_rows():
    from probaforms_b200.models import RealNVP
    rng = np.random.default_rng(2)
    X, Cn = rng.normal(size=(3000, 6)), rng.normal(size=(3000, 2))
    torch.manual_seed(0)
    m = RealNVP(n_layers=4, hidden=(16,), n_epochs=1, batch_size=512)
    m.fit(X, Cn)
    one = m.sample(Cn, seed=5)
    assert one.shape == (3000, 6) and one.dtype == np.float32
    devs = list(range(min(torch.cuda.device_count(), 2))) * (2 if torch.cuda.device_count() < 2 else 1)
    assert np.array_equal(m.sample(Cn, seed=5, devices=devs), one)              # blocks on (possibly) several GPUs
    assert np.array_equal(m.sample(Cn, seed=5, shard=True), one)                # no process group: the whole request
    many = m.sample(Cn, seed=5, n_draws=3)
    assert many.shape == (3, 3000, 6) and np.array_equal(many[0], one) and not np.array_equal(many[1], one)
    lp = m.log_prob_rows(X, Cn)
    ref = m.nf.log_prob_rows(torch.as_tensor(X, dtype=torch.float32).cuda(), torch.as_tensor(Cn, dtype=torch.float32).cuda())
    assert np.array_equal(lp, ref.cpu().numpy())
    assert np.array_equal(m.log_prob_rows(X, Cn, devices=devs), lp)
    big = m.sample(np.repeat(Cn, 200, axis=0), seed=6)                           # > 1 MB: the chunked pinned egress path
    assert big.shape == (600000, 6) and np.isfinite(big).all()
    assert np.array_equal(big[:3000:1][:5], m.sample(np.repeat(Cn, 200, axis=0)[:5], seed=6))


def test_fit_validates_inputs_like_the_reference():
    from probaforms_b200.models import RealNVP
    rng = np.random.default_rng(0)
    X, Cn = rng.normal(size=(100, 4)), rng.normal(size=(100, 2))
    m = RealNVP(n_epochs=1)
    with pytest.raises(ValueError):
        m.fit(X, Cn[:50])                       # TensorDataset: "Size mismatch between tensors"
    m = RealNVP(n_epochs=1)
    m.fit(X, Cn)
    with pytest.raises(ValueError):
        m.fit(X[:, :3], Cn)                     # warm start with another width
    with pytest.raises(ValueError):
        m.fit(X, None)
    with pytest.raises(ValueError):
        m.fit(X, Cn[:99])
    m.fit(X.astype(np.float32), torch.as_tensor(Cn))      # mixed containers are fine


def test_checkpoint_resume_continues_the_adam_trajectory():
    """state_dict of model + optimiser -> fresh objects -> continued fit equals an uninterrupted one (moments and
    step count are imported back into the engine); deepcopy of a fitted model keeps training too."""
    from probaforms_b200.models import RealNVP
    rng = np.random.default_rng(1)
    X, Cn = rng.normal(size=(256, 5)), rng.normal(size=(256, 3))

    def run(n_fits, seed_each):
        torch.manual_seed(0)
        m = RealNVP(n_layers=4, hidden=(8,), n_epochs=2, batch_size=64, lr=1e-2)
        for k in range(n_fits):
            torch.manual_seed(seed_each + k)
            m.fit(X, Cn)
        return m

    full = run(2, 100)
    half = run(1, 100)
    sd_model, sd_opt = copy.deepcopy(half.state_dict()), copy.deepcopy(half.opt.state_dict())
    torch.manual_seed(0)
    fresh = RealNVP(n_layers=4, hidden=(8,), n_epochs=2, batch_size=64, lr=1e-2)
    fresh._model_init(X, Cn)
    fresh.load_state_dict(sd_model)
    fresh.opt.load_state_dict(sd_opt)
    torch.manual_seed(101)
    fresh.fit(X, Cn)
    for (k, a), (_, b) in zip(full.nf.state_dict().items(), fresh.nf.state_dict().items()):
        assert torch.equal(a, b), k
    clone = copy.deepcopy(half)
    torch.manual_seed(101)
    clone.fit(X, Cn)
    for (k, a), (_, b) in zip(full.nf.state_dict().items(), clone.nf.state_dict().items()):
        assert torch.equal(a, b), k


def test_upload_paths_agree():
    """numpy float64, numpy float32, pageable and pinned torch rows all land as the same float32 device rows, including
    the chunked double-buffered path for sets above 1 MB."""
    from probaforms_b200 import _lib
    from probaforms_b200.ingest import upload_resident, rows_to_numpy
    lib = _lib.load()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    A = rng.normal(size=(400001, 9))
    want = torch.as_tensor(A, dtype=torch.float32)
    for src in (A, A.astype(np.float32), want.clone(), want.clone().pin_memory(), torch.as_tensor(A)):
        got = upload_resident(lib, src, dev, chunk_bytes=1 << 20)
        torch.cuda.synchronize()
        assert got.dtype == torch.float32 and torch.equal(got.cpu(), want)
    back = rows_to_numpy(lib, got, chunk_bytes=1 << 20)
    assert back.dtype == np.float32 and np.array_equal(back, want.numpy())


def test_device_shuffle_progressive_upload_trains_and_is_deterministic():
    """shuffle='device' from host arrays with large batches: the shard is uploaded in chunks while the first epoch already
    trains on the chunks that arrived (block-wise shuffle), later epochs use a full device permutation."""
    from probaforms_b200.models import RealNVP
    rng = np.random.default_rng(5)
    n = 60000
    Cn = rng.normal(size=(n, 2))
    X = np.concatenate([Cn * 1.5 - 0.5, rng.normal(size=(n, 2))], axis=1) + 0.2 * rng.normal(size=(n, 4))
    runs = []
    for _ in range(2):
        torch.manual_seed(9)
        m = RealNVP(n_layers=4, hidden=(16,), lr=5e-3, n_epochs=3, batch_size=4096, shuffle='device')
        m.fit(X, Cn)
        runs.append(torch.stack(m.loss_history))
        assert m.h2d_bytes_last_fit == n * 6 * 4
    assert runs[0].shape == (3 * 15,)
    # same batches given the torch seed (the gradient atomics of a multi-CTA step are not bit-reproducible)
    assert torch.allclose(runs[0], runs[1], rtol=1e-4, atol=1e-5)
    assert float(runs[0][-5:].mean()) < float(runs[0][:5].mean()) - 0.3  # it trains
    torch.manual_seed(9)
    ref = RealNVP(n_layers=4, hidden=(16,), lr=5e-3, n_epochs=3, batch_size=4096, shuffle='device', ingest='resident')
    ref.fit(X, Cn)
    assert abs(float(torch.stack(ref.loss_history)[-5:].mean()) - float(runs[0][-5:].mean())) < 0.3


def test_fit_epoch_fused_step_equals_the_two_launch_step():
    """README-sized batches (<= 32 rows) run as ONE launch per step inside rnvp_fit_epoch (fit kernel with the Adam update
    fused behind it); the trajectory must equal the fit_step path (backward launch + Adam launch) bit for bit, and a
    ragged last batch as well as a batch of more than 32 rows (two-launch fallback inside the same call) must work."""
    dev = torch.device("cuda:0")
    shape = (2, 1, 8, (10,))
    n = 100                                                # 3 batches of 32 + one of 4
    g = torch.Generator().manual_seed(0)
    X, Cn = torch.randn(n, 2, generator=g).to(dev), torch.randn(n, 1, generator=g).to(dev)
    perm = torch.randperm(n, generator=g).to(dev)
    for bs in (32, 48):
        nf_a, _ = _flow(shape, 11, dev)
        nf_b, _ = _flow(shape, 11, dev)
        ea, eb = nf_a._fused(), nf_b._fused()
        steps = (n + bs - 1) // bs
        la, lb = torch.zeros(steps, device=dev), torch.zeros(steps, device=dev)
        ea.zero_grads(), eb.zero_grads()
        for _ in range(2):                                 # two epochs: the Adam step counter carries over
            ea.fit_epoch(X, Cn, perm, n, bs, 0.01, 0.0, la)
            for s in range(steps):
                nb = min(bs, n - s * bs)
                eb.fit_step(X, Cn, perm[s * bs:s * bs + nb], nb, nb, 0.01, 0.0, lb[s:s + 1])
        assert torch.equal(la, lb)
        assert torch.equal(ea.flat, eb.flat) and torch.equal(ea.packed, eb.packed)
        assert torch.equal(ea.exp_avg, eb.exp_avg) and torch.equal(ea.exp_avg_sq, eb.exp_avg_sq)
        assert float(ea.gpacked.abs().max()) == 0.0 and ea.adam_steps == eb.adam_steps == 2 * steps


def test_sample_result_lives_in_recycled_pinned_memory_and_equals_the_one_shot_path():
    """Large sample() results are pipelined in row chunks into pinned memory lent by ingest.RESULTS: same values as one
    launch over all rows (noise keyed on the global row index), views keep the buffer alive, dropping the array recycles it."""
    import gc
    from probaforms_b200.models import RealNVP
    import probaforms_b200.ingest as I
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    m = RealNVP(n_layers=4, hidden=(64,), batch_size=4096, n_epochs=1, lr=1e-3)
    torch.manual_seed(0)
    m.fit(rng.standard_normal((8192, 32)), rng.standard_normal((8192, 8)))
    eng = m.nf._fused()
    n = 300_001
    Cs = rng.standard_normal((n, 8)).astype(np.float32)
    got = m.sample(Cs, seed=99)
    assert isinstance(got, np.ndarray) and got.dtype == np.float32 and got.shape == (n, 32)
    assert type(got.base).__name__ == "_LentBuffer"
    want = eng.sample(n, torch.from_numpy(Cs).to(dev), seed=99).cpu().numpy()
    assert np.array_equal(got, want)
    chunked = m._sample_to_host(eng, Cs, 0, n, [99], False, chunk_rows=70_000)     # ragged last chunk
    assert np.array_equal(chunked, want)
    many = m.sample(Cs[:50_000], n_draws=3, seed=5)
    assert many.shape == (3, 50_000, 32)
    for k in range(3):
        assert np.array_equal(many[k], eng.sample(50_000, torch.from_numpy(Cs[:50_000]).to(dev), seed=5 + k).cpu().numpy())
    view = got[1000:2000]
    keep = view.copy()
    free0 = I.RESULTS.free_buffers()
    del got, chunked, many
    gc.collect()
    assert I.RESULTS.free_buffers() == free0 + 2                # `view` still pins the first buffer
    again = m.sample(Cs, seed=123)                          # must not overwrite memory a live view points into
    assert np.array_equal(view, keep)
    del view, again
    gc.collect()
    assert I.RESULTS.free_buffers() >= free0 + 2
