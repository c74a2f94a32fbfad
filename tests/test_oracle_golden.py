"""Pin the oracle (oracle/realnvp_oracle.py) to vectors produced by the real reference.

CPU only.  The fixtures come from tests/golden/make_golden.py (reference run
unmodified on CPU).  The oracle uses the same ATen ops in the same order, so
outputs are required to be bit-identical, not merely close.
"""
import numpy as np
import pytest
import torch

from conftest import SMALL_CASES, SEEDED_CASES, FIT_CASES, load_golden, golden_params
from oracle import realnvp_oracle as O


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)       # fixtures were generated single-threaded (summation order)
    yield
    torch.set_num_threads(n)


def _cfg(g):
    hidden = tuple(int(h) for h in g["hidden"])
    return int(g["D"]), int(g["Cd"]), int(g["L"]), hidden, str(g["activation"])


def _xc(g):
    X = torch.from_numpy(g["X"])
    C = torch.from_numpy(g["C"]) if "C" in g else None
    return X, C


@pytest.mark.parametrize("name", SMALL_CASES)
def test_init_matches_reference_rng_order(name):
    g = load_golden(name)
    D, Cd, L, hidden, act = _cfg(g)
    p = O.init_params(D, Cd, L, hidden, seed=int(g["seed"]))
    ref = golden_params(g)
    assert list(p.keys()) == O.param_order(L, len(hidden))
    assert set(p.keys()) == set(ref.keys())
    for k in ref:
        assert torch.equal(p[k], ref[k]), k


@pytest.mark.parametrize("name", SMALL_CASES)
def test_forward_inverse_bit_exact(name):
    g = load_golden(name)
    D, Cd, L, hidden, act = _cfg(g)
    p = golden_params(g)
    X, C = _xc(g)
    z, ld, lp = O.flow_forward_rows(X, C, p, L, len(hidden), act)
    assert np.array_equal(z.numpy(), g["z"])
    assert np.array_equal(ld.numpy(), g["logdet"])
    assert np.array_equal(lp.numpy(), g["logp"])
    assert np.array_equal(O.flow_log_prob(X, C, p, L, len(hidden), act).numpy(), g["log_prob_mean"])
    s = O.flow_sample_from_noise(torch.from_numpy(g["eps"]), C, p, L, len(hidden), act)
    assert np.array_equal(s.numpy(), g["sample"])
    y1, ld1 = O.coupling_f(X, C, p, 1, len(hidden), act)
    assert np.array_equal(y1.numpy(), g["layer1_f"])
    assert np.array_equal(ld1.numpy(), g["layer1_logdet"])
    assert np.array_equal(O.coupling_g(X, C, p, 1, len(hidden), act).numpy(), g["layer1_g"])


@pytest.mark.parametrize("name", SMALL_CASES)
def test_gradients_bit_exact_and_masked_zeros(name):
    g = load_golden(name)
    D, Cd, L, hidden, act = _cfg(g)
    p = golden_params(g)
    X, C = _xc(g)
    loss, grads = O.loss_and_grads(X, C, p, L, len(hidden), act)
    assert np.array_equal(loss.numpy(), g["loss"])
    for k, gr in grads.items():
        assert np.array_equal(gr.numpy(), g["g/" + k]), k
    # W1 columns of the transformed set and W2 rows / b2 entries of the kept set are exactly 0
    nh = len(hidden)
    for i in range(L):
        mask = O.layer_mask(D, i).numpy()
        for net in "ts":
            w1 = grads[f"layers.{i}.nn_{net}.0.weight"].numpy()
            assert np.all(w1[:, :D][:, mask == 0] == 0.0)
            w2 = grads[f"layers.{i}.nn_{net}.{2 * nh}.weight"].numpy()
            b2 = grads[f"layers.{i}.nn_{net}.{2 * nh}.bias"].numpy()
            assert np.all(w2[mask == 1] == 0.0) and np.all(b2[mask == 1] == 0.0)


@pytest.mark.parametrize("name", SMALL_CASES)
@pytest.mark.parametrize("tag,wd", [("adam0", 0.0), ("adamwd", 0.2)])
def test_adam_restatement(name, tag, wd):
    g = load_golden(name)
    D, Cd, L, hidden, act = _cfg(g)
    p = golden_params(g)
    X, C = _xc(g)
    st = O.AdamState(p, lr=0.01, weight_decay=wd)
    losses = []
    for _ in range(3):
        loss, grads = O.loss_and_grads(X, C, p, L, len(hidden), act)
        O.adam_step(p, grads, st)
        losses.append(loss.numpy())
    assert np.array_equal(np.array(losses), g[tag + "/losses"])
    for k in p:
        assert np.array_equal(p[k].numpy(), g[tag + "/" + k]), k


@pytest.mark.parametrize("name", SEEDED_CASES)
def test_seeded_bench_shapes(name):
    g = load_golden(name)
    D, Cd, L, hidden, act = _cfg(g)
    seed, N = int(g["seed"]), int(g["N"])
    p = O.init_params(D, Cd, L, hidden, seed=seed)
    flat = torch.cat([p[k].reshape(-1) for k in O.param_order(L, len(hidden))])
    idx = torch.from_numpy(g["grad_idx"])
    assert np.array_equal(flat[idx].numpy(), g["param_idx_vals"])
    assert float(flat.double().sum()) == float(g["param_sum"])
    gen = torch.Generator().manual_seed(seed + 1000)
    X = torch.randn(N, D, generator=gen)
    C = torch.randn(N, Cd, generator=gen)
    eps = torch.randn(N, D, generator=gen)
    z, ld, lp = O.flow_forward_rows(X, C, p, L, len(hidden), act)
    assert np.array_equal(z.numpy(), g["z"])
    assert np.array_equal(lp.numpy(), g["logp"])
    s = O.flow_sample_from_noise(eps, C, p, L, len(hidden), act)
    assert np.array_equal(s.numpy(), g["sample"])
    loss, grads = O.loss_and_grads(X, C, p, L, len(hidden), act)
    gflat = torch.cat([grads[k].reshape(-1) for k in O.param_order(L, len(hidden))])
    assert np.array_equal(gflat[idx].numpy(), g["grad_vals"])
    assert int((gflat != 0).sum()) == int(g["grad_nnz"])


@pytest.mark.parametrize("name", FIT_CASES)
def test_fit_loop_restatement(name):
    """DataLoader RNG consumption, batching, Adam: whole-trajectory equality with the reference."""
    g = load_golden(name)
    X = g["X"]
    C = g["C"] if "C" in g else None
    seed = int(g["seed"])
    D = X.shape[1]
    Cd = 0 if C is None else C.shape[1]
    torch.manual_seed(seed)
    p = O.init_params(D, Cd, 8, (10,))                 # RealNVP defaults, realnvp.py:160
    hist, _ = O.fit(X, C, p, 8, 1, "tanh", batch_size=32, n_epochs=int(g["n_epochs"]), lr=0.01,
                    weight_decay=float(g["weight_decay"]))
    assert np.array_equal(np.array([float(h) for h in hist], dtype=np.float32), g["loss_history"])
    for k in p:
        assert np.array_equal(p[k].numpy(), g["p/" + k]), k
    torch.manual_seed(seed + 1)
    s = O.flow_sample(torch.as_tensor(C, dtype=torch.float32) if C is not None else X.shape[0],
                      D, p, 8, 1, "tanh")
    assert np.array_equal(s.numpy(), g["sample"])
