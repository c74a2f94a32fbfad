"""GPU parity tests (run on the B200 box with ``-m gpu``).

Every test drives the CUDA path through the reference-shaped Python API, i.e. through the C ABI
of librnvp_b200.so, and compares with (a) the committed golden vectors the real reference produced
and (b) the CPU oracle on the same seeded inputs.  Tolerance: north_star's fp32 bound, rel 1e-5
measured as max-abs error / max-abs reference value; the reference's own fp32-vs-fp64
noise floor is 3e-7..2.6e-6.
"""
import copy

import numpy as np
import pytest
import torch

from conftest import SMALL_CASES, SEEDED_CASES, FIT_CASES, load_golden, golden_params, rel_err
from oracle import realnvp_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-5
GRAD_TOL = 2e-5       # gradients: atomically accumulated over row tiles (order varies run to run)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def build_flow(D, Cd, L, hidden, act, params, dev):
    from probaforms_b200.models import RealNVPLayer, NormalizingFlow
    layers = [RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, hidden, act) for i in range(L)]
    nf = NormalizingFlow(layers, prior=None)
    if params is not None:
        nf.load_state_dict(params)
    return nf.to(dev)


def cfg(g):
    return int(g["D"]), int(g["Cd"]), int(g["L"]), tuple(int(h) for h in g["hidden"]), str(g["activation"])


def xc(g, dev):
    X = torch.from_numpy(g["X"]).to(dev)
    C = torch.from_numpy(g["C"]).to(dev) if "C" in g else None
    return X, C


def test_native_library_is_what_runs(dev):
    """The product path must be the CUDA extension: loaded from the in-tree .so, no fallback."""
    from probaforms_b200 import _lib
    lib = _lib.load()
    assert lib.rnvp_version() >= 100
    with open("/proc/self/maps") as f:
        assert "librnvp_b200.so" in f.read()


@pytest.mark.parametrize("name", SMALL_CASES)
def test_forward_inverse_layer_parity(name, dev):
    g = load_golden(name)
    D, Cd, L, hidden, act = cfg(g)
    nf = build_flow(D, Cd, L, hidden, act, golden_params(g), dev)
    X, C = xc(g, dev)
    z, ld, lp = nf.forward_rows(X, C)
    assert rel_err(z.cpu(), g["z"]) < TOL
    assert rel_err(ld.cpu(), g["logdet"]) < TOL
    assert rel_err(lp.cpu(), g["logp"]) < TOL
    assert rel_err(nf.log_prob_rows(X, C).cpu(), g["logp"]) < TOL
    with torch.no_grad():
        assert abs(float(nf.log_prob(X, C)) - float(g["log_prob_mean"])) < TOL * abs(float(g["log_prob_mean"]))
    eps = torch.from_numpy(g["eps"]).to(dev)
    assert rel_err(nf.sample_from_noise(eps, C).cpu(), g["sample"]) < TOL
    if L > 1:
        y1, ld1 = nf.layers[1].f(X, C)
        assert rel_err(y1.cpu(), g["layer1_f"]) < TOL
        assert np.max(np.abs(ld1.cpu().numpy() - g["layer1_logdet"])) < TOL * max(1.0, np.abs(g["layer1_logdet"]).max())
        assert rel_err(nf.layers[1].g(X, C).cpu(), g["layer1_g"]) < TOL


@pytest.mark.parametrize("name", SMALL_CASES)
def test_gradient_parity_and_exact_zeros(name, dev):
    g = load_golden(name)
    D, Cd, L, hidden, act = cfg(g)
    nf = build_flow(D, Cd, L, hidden, act, golden_params(g), dev)
    X, C = xc(g, dev)
    loss = -nf.log_prob(X, C)
    assert loss.dim() == 0
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < TOL * abs(float(g["loss"]))
    gmax = max(np.abs(g["g/" + k]).max() for k, _ in nf.named_parameters())
    for k, p in nf.named_parameters():
        ref = g["g/" + k]
        assert p.grad is not None and p.grad.shape == p.shape
        got = p.grad.cpu().numpy()
        assert np.max(np.abs(got - ref)) < GRAD_TOL * gmax, k
    # masked rows / columns: exactly 0.0, tensors present
    nh = len(hidden)
    for i in range(L):
        mask = ((np.arange(D) + i) % 2)
        for net in "ts":
            w1 = dict(nf.named_parameters())[f"layers.{i}.nn_{net}.0.weight"].grad.cpu().numpy()
            assert np.all(w1[:, :D][:, mask == 0] == 0.0)
            w2 = dict(nf.named_parameters())[f"layers.{i}.nn_{net}.{2 * nh}.weight"].grad.cpu().numpy()
            b2 = dict(nf.named_parameters())[f"layers.{i}.nn_{net}.{2 * nh}.bias"].grad.cpu().numpy()
            assert np.all(w2[mask == 1] == 0.0) and np.all(b2[mask == 1] == 0.0)


@pytest.mark.parametrize("name", ["t5c3_tanh", "multi_hidden_relu", "d1_regression", "nocond_d5"])
@pytest.mark.parametrize("tag,wd", [("adam0", 0.0), ("adamwd", 0.2)])
def test_adam_steps_parity(name, tag, wd, dev):
    from probaforms_b200.models.realnvp import FusedAdam
    g = load_golden(name)
    D, Cd, L, hidden, act = cfg(g)
    nf = build_flow(D, Cd, L, hidden, act, golden_params(g), dev)
    X, C = xc(g, dev)
    opt = FusedAdam(nf, lr=0.01, weight_decay=wd)
    losses = []
    for _ in range(3):
        loss = -nf.log_prob(X, C)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert np.allclose(losses, g[tag + "/losses"], rtol=1e-5, atol=1e-6)
    for k, p in nf.state_dict().items():
        ref = g[tag + "/" + k]
        # after 3 steps of lr=0.01 every entry moved by ~0.03: compare the update, not just the value
        assert np.max(np.abs(p.cpu().numpy() - ref)) < 2e-5 * max(1.0, np.abs(ref).max()), k
    st = opt.state_dict()
    assert len(st["state"]) == len(list(nf.parameters()))


@pytest.mark.parametrize("name", SEEDED_CASES)
def test_bench_shapes_against_reference_goldens(name, dev):
    g = load_golden(name)
    D, Cd, L, hidden, act = cfg(g)
    seed, N = int(g["seed"]), int(g["N"])
    torch.manual_seed(seed)
    nf = build_flow(D, Cd, L, hidden, act, None, dev)        # default init in the reference's RNG order
    flat = torch.cat([p.detach().reshape(-1) for p in nf.parameters()]).cpu()
    idx = torch.from_numpy(g["grad_idx"])
    assert np.array_equal(flat[idx].numpy(), g["param_idx_vals"])
    gen = torch.Generator().manual_seed(seed + 1000)
    X = torch.randn(N, D, generator=gen).to(dev)
    C = torch.randn(N, Cd, generator=gen).to(dev)
    eps = torch.randn(N, D, generator=gen).to(dev)
    z, ld, lp = nf.forward_rows(X, C)
    assert rel_err(z.cpu(), g["z"]) < TOL and rel_err(lp.cpu(), g["logp"]) < TOL
    assert rel_err(nf.sample_from_noise(eps, C).cpu(), g["sample"]) < TOL
    loss = -nf.log_prob(X, C)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < TOL * abs(float(g["loss"]))
    gflat = torch.cat([p.grad.reshape(-1) for p in nf.parameters()]).cpu()
    err = float((gflat[idx] - torch.from_numpy(g["grad_vals"])).abs().max()) / float(g["grad_absmax"])
    assert err < GRAD_TOL, err
    assert abs(float(gflat.double().norm()) - float(g["grad_l2"])) < 1e-4 * float(g["grad_l2"])
    assert int((gflat != 0).sum()) <= int(g["grad_nnz"])


@pytest.mark.parametrize("name", FIT_CASES)
def test_fit_trajectory_matches_reference(name, dev):
    """RealNVP.fit end to end: same init, same epoch permutations, same losses and weights."""
    from probaforms_b200.models import RealNVP
    g = load_golden(name)
    X = g["X"]
    C = g["C"] if "C" in g else None
    seed = int(g["seed"])
    torch.manual_seed(seed)
    model = RealNVP(lr=0.01, n_epochs=int(g["n_epochs"]), weight_decay=float(g["weight_decay"]))
    assert model.fit(X, C) is None
    hist = np.array([float(l) for l in model.loss_history], dtype=np.float32)
    assert all(l.dim() == 0 and l.device.type == "cpu" for l in model.loss_history)
    assert hist.shape == g["loss_history"].shape
    assert np.allclose(hist, g["loss_history"], rtol=2e-5, atol=2e-6), np.abs(hist - g["loss_history"]).max()
    for k, v in model.nf.state_dict().items():
        assert np.max(np.abs(v.cpu().numpy() - g["p/" + k])) < 5e-5 * max(1.0, np.abs(g["p/" + k]).max()), k
    assert set("nf." + k for k in model.nf.state_dict()) == set(model.state_dict().keys())
    # sample from the noise the reference drew (prior.sample == randn on the CPU default generator)
    torch.manual_seed(seed + 1)
    n = X.shape[0]
    eps = torch.randn(n, X.shape[1])
    Cd = None if C is None else torch.as_tensor(C, dtype=torch.float32, device=dev)
    s = model.nf.sample_from_noise(eps.to(dev), Cd)
    assert np.max(np.abs(s.cpu().numpy() - g["sample"])) < 1e-3 * max(1.0, np.abs(g["sample"]).max())
    out = model.sample(C if C is not None else n)
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == X.shape
    # warm start: a second fit reuses nf / opt and appends to loss_history
    nf_before, n_hist = model.nf, len(model.loss_history)
    model.fit(X, C)
    assert model.nf is nf_before and len(model.loss_history) == 2 * n_hist


def test_upstream_smoke_tests(dev):
    """The reference's own two tests (tests/test_models.py:11-28), unchanged in spirit."""
    from probaforms_b200.models import RealNVP, GenModel
    assert issubclass(RealNVP, GenModel) and issubclass(RealNVP, torch.nn.Module)
    n = 100
    X = np.random.normal(size=(n, 5))
    C = np.random.normal(size=(n, 3))
    gen = RealNVP()
    gen.fit(X, C)
    assert gen.sample(C).shape == X.shape
    gen = RealNVP()
    gen.fit(X, C=None)
    assert gen.sample(C=n).shape == X.shape
    for attr in ("nf", "opt", "prior", "loss_history", "n_layers", "hidden", "activation", "batch_size",
                 "n_epochs", "lr", "weight_decay", "verbose"):
        assert hasattr(gen, attr)
    assert len(gen.loss_history) == 10 * 4


@pytest.mark.parametrize("shape", [(32, 8, 16, (128,)), (2, 1, 8, (10,)), (64, 16, 4, (128,)), (7, 0, 5, (16, 12))])
@pytest.mark.parametrize("N", [1, 63, 65, 1000, 70001])
def test_roundtrip_and_ragged_sizes(shape, N, dev):
    """g(f(x)) == x, tails that do not fill a tile, and oracle agreement at each size."""
    D, Cd, L, hidden = shape
    torch.manual_seed(5)
    nf = build_flow(D, Cd, L, hidden, "tanh", None, dev)
    gen = torch.Generator().manual_seed(N)
    X = torch.randn(N, D, generator=gen)
    C = torch.randn(N, Cd, generator=gen) if Cd else None
    Xd, Cdv = X.to(dev), (C.to(dev) if Cd else None)
    z, ld, lp = nf.forward_rows(Xd, Cdv)
    back = nf.sample_from_noise(z, Cdv)
    assert float((back - Xd).abs().max()) < 1e-4 * max(1.0, float(Xd.abs().max()))
    if N <= 1000:
        params = {k: v.detach().cpu() for k, v in nf.state_dict().items()}
        zr, ldr, lpr = O.flow_forward_rows(X, C, params, L, len(hidden), "tanh")
        assert rel_err(z.cpu(), zr) < TOL and rel_err(lp.cpu(), lpr) < TOL


def test_gradient_is_additive_over_row_shards(dev):
    """Size-independent property used for data parallelism: grad(sum over rows) = sum of shard grads,
    and the row gather (epoch permutation slice) equals materialising the rows."""
    torch.manual_seed(3)
    D, Cd, L, hidden = 32, 8, 16, (128,)
    nf = build_flow(D, Cd, L, hidden, "tanh", None, dev)
    eng = nf._fused()
    N = 40000
    X = torch.randn(N, D, device=dev)
    C = torch.randn(N, Cd, device=dev)
    eng.zero_grads()
    eng.backward(X, C, None, N, -1.0 / N)
    g_all = eng.unpack_grads().clone()
    loss_all = float(eng.loss_slot)
    eng.zero_grads()
    h = 17001
    eng.backward(X[:h].contiguous(), C[:h].contiguous(), None, h, -1.0 / N)
    eng.backward(X[h:].contiguous(), C[h:].contiguous(), None, N - h, -1.0 / N)
    g_two = eng.unpack_grads().clone()
    assert abs(float(eng.loss_slot) - loss_all) < 1e-4 * abs(loss_all)
    scale = float(g_all.abs().max())
    assert float((g_all - g_two).abs().max()) < 2e-5 * scale
    perm = torch.randperm(N, device=dev)
    eng.zero_grads()
    eng.backward(X, C, perm, N, -1.0 / N)
    g_perm = eng.unpack_grads().clone()
    assert float((g_all - g_perm).abs().max()) < 2e-5 * scale
    eng.zero_grads()
    # log-prob rows from the fit-step kernels == forward kernel of the same family (bit-equal), and the two
    # families (FP32-FMA tile kernels / tcgen05 TF32x3 kernels) agree within the fp32 tolerance
    lps = {}
    for path in (1, 0):
        eng.set_path(path)
        lp_b = torch.empty(N, device=dev)
        eng.backward(X, C, None, N, -1.0 / N, logp_rows=lp_b)
        lp_f = eng.forward(X, C, want_z=False, want_logdet=False)[2]
        assert torch.equal(lp_b, lp_f)
        lps[path] = (lp_f, eng.unpack_grads().clone())
        eng.zero_grads()
    assert float((lps[0][0] - lps[1][0]).abs().max()) < 1e-5 * float(lps[1][0].abs().max())
    assert float((lps[0][1] - lps[1][1]).abs().max()) < 2e-5 * float(lps[1][1].abs().max())


def test_error_behaviour(dev):
    from probaforms_b200.models import RealNVP
    from probaforms_b200._lib import RnvpError
    torch.manual_seed(0)
    nf = build_flow(5, 3, 4, (10,), "tanh", None, dev)
    X = torch.randn(10, 5, device=dev)
    with pytest.raises(RuntimeError):
        nf.forward_rows(X, None)                       # conditions missing
    with pytest.raises(RuntimeError):
        nf.forward_rows(torch.randn(10, 4, device=dev), torch.randn(10, 3, device=dev))   # wrong width
    cpu_flow = build_flow(5, 3, 4, (10,), "tanh", None, torch.device("cpu"))
    with pytest.raises(RuntimeError):
        cpu_flow.forward_rows(torch.randn(4, 5), torch.randn(4, 3))          # no CPU path
    eng = nf._fused()
    with pytest.raises(RnvpError):
        eng.forward(X, torch.randn(10, 3, device=dev), layers=(3, 9))        # bad layer range
    m = RealNVP()
    with pytest.raises(TypeError):
        m.fit(np.zeros((4, 2)), None)
        m.sample(np.int64(5))       # numpy ints are not the int path upstream either (len() of a 0-d)
    # empty input
    z, ld, lp = nf.forward_rows(torch.empty(0, 5, device=dev), torch.empty(0, 3, device=dev))
    assert z.shape == (0, 5) and lp.shape == (0,)


def test_state_dict_roundtrip_and_deepcopy(dev):
    torch.manual_seed(1)
    nf = build_flow(6, 2, 4, (10, 20, 15), "relu", None, dev)
    X, C = torch.randn(50, 6, device=dev), torch.randn(50, 2, device=dev)
    lp = nf.log_prob_rows(X, C)
    sd = {k: v.clone() for k, v in nf.state_dict().items()}
    nf2 = build_flow(6, 2, 4, (10, 20, 15), "relu", None, dev)
    nf2.load_state_dict(sd)
    assert torch.equal(nf2.log_prob_rows(X, C), lp)
    nf3 = copy.deepcopy(nf)
    assert torch.equal(nf3.log_prob_rows(X, C), lp)
    with torch.no_grad():
        for p in nf3.parameters():
            p.mul_(1.5)                                # in-place edits are picked up (re-packed) ...
    assert not torch.equal(nf3.log_prob_rows(X, C), lp)
    assert torch.equal(nf.log_prob_rows(X, C), lp)     # ... and the copy does not alias the original


@pytest.mark.gpu
def test_repeated_sampling_equals_a_loop_of_sample_calls(dev):
    """sample(C, n_draws=k) (the notebooks' Monte-Carlo loop) == k calls of sample(C) under the same seed."""
    from probaforms_b200.models import RealNVP
    rng = np.random.default_rng(3)
    X = rng.normal(size=(200, 5))
    C = rng.normal(size=(200, 3))
    torch.manual_seed(0)
    gen = RealNVP(n_epochs=2)
    gen.fit(X, C)
    torch.manual_seed(5)
    torch.cuda.manual_seed(5)
    want = np.stack([gen.sample(C) for _ in range(4)])
    torch.manual_seed(5)
    torch.cuda.manual_seed(5)
    got = gen.sample(C, n_draws=4)
    assert got.shape == (4, 200, 5) and got.dtype == np.float32
    assert np.array_equal(got, want)
    gen2 = RealNVP(n_epochs=1)
    gen2.fit(X, None)
    assert gen2.sample(7, n_draws=3).shape == (3, 7, 5)


@pytest.mark.gpu
def test_device_shuffle_option_trains_and_is_reproducible(dev):
    """shuffle='device' (opt-in): same per-epoch RNG consumption, GPU randperm instead of the reference's CPU order."""
    from probaforms_b200.models import RealNVP
    rng = np.random.default_rng(1)
    C = rng.normal(size=(512, 2))
    X = np.concatenate([C * 2.0 + 1.0, rng.normal(size=(512, 2))], axis=1) + 0.1 * rng.normal(size=(512, 4))
    runs = []
    for _ in range(2):
        torch.manual_seed(3)
        gen = RealNVP(n_layers=4, hidden=(16,), lr=5e-3, n_epochs=15, batch_size=64, shuffle='device')
        gen.fit(X, C)
        runs.append(torch.stack(gen.loss_history))
        after = torch.rand(1)
    assert torch.equal(runs[0], runs[1])                              # deterministic given the torch seed
    assert float(runs[0][-8:].mean()) < float(runs[0][:8].mean()) - 0.5   # it trains
    torch.manual_seed(3)
    ref = RealNVP(n_layers=4, hidden=(16,), lr=5e-3, n_epochs=15, batch_size=64)
    ref.fit(X, C)
    assert torch.equal(after, torch.rand(1))                          # both modes consume the global RNG identically
    assert abs(float(torch.stack(ref.loss_history)[-8:].mean()) - float(runs[0][-8:].mean())) < 0.5
    with pytest.raises(ValueError):
        RealNVP(shuffle='nope')
