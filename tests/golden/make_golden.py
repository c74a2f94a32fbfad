"""Generate golden vectors by running the REAL reference (hse-cs/probaforms).

Run in the build container only (the reference is mounted read-only at
/root/reference and does not exist on the GPU box):

    python tests/golden/make_golden.py

Every ``*.npz`` written next to this file holds inputs and the outputs the
unmodified reference produced for them on CPU (torch 2.11.0, fp32).  The
fixtures pin ``oracle/realnvp_oracle.py`` (tests/test_oracle_golden.py) and,
through it and directly, the CUDA path (tests/test_gpu_parity.py).

Small cases store the parameters themselves.  The three bench-shaped cases
(c3/c4/c5) would be megabytes of weights, so they store only the seed: the
reference draws its default nn.Linear init from the global torch RNG in a fixed
order (realnvp.py:69-70), which ``oracle.init_params(seed=...)`` reproduces --
the small cases prove that equivalence bit for bit.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("PROBAFORMS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
os.environ.pop("device", None)          # reference: env var 'device' unset -> CPU (realnvp.py:12-15)

from probaforms.models import RealNVP                       # noqa: E402
from probaforms.models.realnvp import RealNVPLayer          # noqa: E402
from probaforms.models.nflow import NormalizingFlow         # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)                # deterministic summation order for the fixtures


def build_reference_flow(D, Cd, L, hidden, activation, seed):
    """Exactly RealNVP._model_init (realnvp.py:180-207) minus the optimiser."""
    torch.manual_seed(seed)
    prior = torch.distributions.MultivariateNormal(torch.zeros(D), torch.eye(D))
    layers = [RealNVPLayer(var_size=D, cond_size=Cd, mask=((torch.arange(D) + i) % 2),
                           hidden=hidden, activation=activation) for i in range(L)]
    return NormalizingFlow(layers=layers, prior=prior)


def rows(nf, X, C):
    """nflow.py:107-115 without the mean."""
    ll = None
    for layer in nf.layers:
        X, change = layer.f(X, C)
        ll = change if ll is None else ll + change
    return X, ll, ll + nf.prior.log_prob(X)


def sample_from(nf, eps, C):
    X = eps
    for layer in nf.layers[::-1]:
        X = layer.g(X, C)
    return X


def data(D, Cd, N, seed):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(N, D, generator=g)
    C = torch.randn(N, Cd, generator=g) if Cd > 0 else None
    eps = torch.randn(N, D, generator=g)
    return X, C, eps


def small_case(name, D, Cd, L, hidden, activation, N, seed):
    nf = build_reference_flow(D, Cd, L, hidden, activation, seed)
    X, C, eps = data(D, Cd, N, seed + 1000)
    out = {"D": D, "Cd": Cd, "L": L, "hidden": np.array(hidden), "activation": activation,
           "seed": seed, "X": X.numpy(), "eps": eps.numpy()}
    if C is not None:
        out["C"] = C.numpy()
    sd0 = {k: v.detach().clone() for k, v in nf.state_dict().items()}
    for k, v in sd0.items():
        out["p/" + k] = v.numpy()
    with torch.no_grad():
        z, ld, lp = rows(nf, X, C)
        out["z"], out["logdet"], out["logp"] = z.numpy(), ld.numpy(), lp.numpy()
        out["log_prob_mean"] = nf.log_prob(X, C).numpy()
        out["sample"] = sample_from(nf, eps, C).numpy()
        # per-layer f / g of layer 1 (odd mask) for the layer-level API
        y1, ld1 = nf.layers[1].f(X, C)
        out["layer1_f"], out["layer1_logdet"] = y1.numpy(), ld1.numpy()
        out["layer1_g"] = nf.layers[1].g(X, C).numpy()
    # gradients of loss = -log_prob (realnvp.py:246-250)
    loss = -nf.log_prob(X, C)
    nf.zero_grad()
    loss.backward()
    out["loss"] = loss.detach().numpy()
    for k, p in nf.named_parameters():
        out["g/" + k] = p.grad.detach().numpy().copy()
    # k Adam steps on the same batch, wd = 0 and wd = 0.2 (forecast.ipynb cell 23)
    for tag, wd in (("adam0", 0.0), ("adamwd", 0.2)):
        nf.load_state_dict(sd0)
        opt = torch.optim.Adam(nf.parameters(), lr=0.01, weight_decay=wd)
        losses = []
        for _ in range(3):
            loss = -nf.log_prob(X, C)
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(loss.detach().numpy())
        out[tag + "/losses"] = np.array(losses)
        for k, v in nf.state_dict().items():
            out[tag + "/" + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "P =", sum(v.numel() for v in sd0.values()))


def seeded_case(name, D, Cd, L, hidden, activation, N, seed, n_grad_samples=4096):
    """Bench-shaped flows: store the seed instead of the weights."""
    nf = build_reference_flow(D, Cd, L, hidden, activation, seed)
    X, C, eps = data(D, Cd, N, seed + 1000)
    out = {"D": D, "Cd": Cd, "L": L, "hidden": np.array(hidden), "activation": activation,
           "seed": seed, "N": N}
    with torch.no_grad():
        z, ld, lp = rows(nf, X, C)
        out["z"], out["logdet"], out["logp"] = z.numpy(), ld.numpy(), lp.numpy()
        out["sample"] = sample_from(nf, eps, C).numpy()
    loss = -nf.log_prob(X, C)
    loss.backward()
    out["loss"] = loss.detach().numpy()
    flat = torch.cat([p.grad.reshape(-1) for p in nf.parameters()])
    pflat = torch.cat([p.detach().reshape(-1) for p in nf.parameters()])
    gi = torch.Generator().manual_seed(7)
    idx = torch.randperm(flat.numel(), generator=gi)[:n_grad_samples]
    out["grad_idx"] = idx.numpy()
    out["grad_vals"] = flat[idx].numpy()
    out["grad_absmax"] = flat.abs().max().numpy()
    out["grad_l2"] = flat.double().norm().numpy()
    out["grad_nnz"] = int((flat != 0).sum())
    out["param_sum"] = pflat.double().sum().numpy()          # pins init_params(seed) equivalence
    out["param_idx_vals"] = pflat[idx].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "P =", flat.numel())


def fit_case(name, N, n_epochs, seed, with_cond=True, weight_decay=0.0):
    """End-to-end RealNVP.fit / .sample through the reference's public API."""
    rng = np.random.RandomState(seed)
    # two interleaving half circles, the README's make_moons recipe without sklearn's shuffle
    n0 = N // 2
    th0, th1 = rng.uniform(0, np.pi, n0), rng.uniform(0, np.pi, N - n0)
    X = np.concatenate([np.stack([np.cos(th0), np.sin(th0)], 1),
                        np.stack([1 - np.cos(th1), 0.5 - np.sin(th1)], 1)])
    X = X + 0.1 * rng.normal(size=X.shape)
    y = np.concatenate([np.zeros(n0), np.ones(N - n0)])
    C = y.reshape(-1, 1) if with_cond else None
    torch.manual_seed(seed)
    model = RealNVP(lr=0.01, n_epochs=n_epochs, weight_decay=weight_decay)
    model.fit(X, C)
    out = {"X": X, "seed": seed, "n_epochs": n_epochs, "weight_decay": weight_decay,
           "loss_history": np.array([float(l) for l in model.loss_history], dtype=np.float32)}
    if with_cond:
        out["C"] = C
    for k, v in model.nf.state_dict().items():
        out["p/" + k] = v.numpy().copy()
    torch.manual_seed(seed + 1)
    out["sample"] = model.sample(C if with_cond else N)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "steps =", len(model.loss_history), "final loss", out["loss_history"][-1])


if __name__ == "__main__":
    small_case("t5c3_tanh", D=5, Cd=3, L=8, hidden=(10,), activation="tanh", N=64, seed=0)
    small_case("moons_shape", D=2, Cd=1, L=8, hidden=(10,), activation="tanh", N=128, seed=1)
    small_case("d1_regression", D=1, Cd=1, L=4, hidden=(10,), activation="tanh", N=50, seed=2)
    small_case("nocond_d5", D=5, Cd=0, L=8, hidden=(10,), activation="tanh", N=37, seed=3)
    small_case("multi_hidden_relu", D=6, Cd=2, L=4, hidden=(10, 20, 15), activation="relu", N=70, seed=4)
    small_case("unknown_act", D=4, Cd=2, L=3, hidden=(8,), activation="sigmoid", N=33, seed=5)
    small_case("multi_hidden_tanh", D=7, Cd=3, L=5, hidden=(12, 9), activation="tanh", N=65, seed=6)
    seeded_case("c3_shape", D=32, Cd=8, L=16, hidden=(128,), activation="tanh", N=192, seed=10)
    seeded_case("c4_shape", D=64, Cd=16, L=24, hidden=(128,), activation="tanh", N=96, seed=11)
    seeded_case("c5_shape", D=128, Cd=32, L=8, hidden=(512,), activation="tanh", N=80, seed=12)
    fit_case("fit_moons", N=100, n_epochs=2, seed=0, with_cond=True)
    fit_case("fit_nocond_wd", N=70, n_epochs=2, seed=3, with_cond=False, weight_decay=0.2)
