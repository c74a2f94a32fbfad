"""CPU oracle for the probaforms RealNVP hot path.  TEST INFRASTRUCTURE ONLY.

This module is a functional restatement (torch CPU ops, no nn.Module machinery)
of what hse-cs/probaforms computes on the RealNVP path.  It exists so that the
CUDA product path can be checked against something that is *not* the product.
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it; nothing under
``probaforms_b200/`` does.

Where the arithmetic lives: the reference's maths is executed by third-party
PyTorch (``torch = "^2.0.0"`` in the reference's pyproject.toml:18, no lock
file; this image has torch 2.11.0+cu128).  The restatement therefore uses the
same ATen ops in the same order as the reference call sites cited below, so on
CPU it is bit-identical to the reference (pinned by
``tests/golden/make_golden.py`` -> ``tests/test_oracle_golden.py``; the
reference's own tests hold no golden vectors).

Parameter container: a plain ``dict`` with the reference's ``state_dict`` key
names relative to the flow, ``layers.{i}.nn_{t,s}.{2k}.{weight,bias}``
(reference probaforms/models/realnvp.py:69-70 + nflow.py:84).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# construction (reference realnvp.py:19-43, 65-70, 195-202)
# --------------------------------------------------------------------------
def layer_mask(var_size: int, i: int) -> torch.Tensor:
    """mask of layer i: (arange(D) + i) % 2   -- realnvp.py:199."""
    return (torch.arange(var_size) + i) % 2


def is_tanh(activation: str) -> bool:
    """'tanh' -> Tanh; 'relu' *and any other string* -> ReLU -- realnvp.py:32-37."""
    return activation == "tanh"


def net_keys(i: int, net: str, n_hidden: int) -> List[Tuple[str, str]]:
    """(weight key, bias key) for every Linear of nn_{net} of layer i, in order."""
    return [(f"layers.{i}.nn_{net}.{2 * q}.weight", f"layers.{i}.nn_{net}.{2 * q}.bias")
            for q in range(n_hidden + 1)]


def init_params(var_size: int, cond_size: int, n_layers: int, hidden: Sequence[int],
                seed: Optional[int] = None, dtype=torch.float32) -> Params:
    """Default-initialised parameters drawn in the reference's RNG order.

    The reference builds, per layer, nn_t first and then nn_s (realnvp.py:69-70),
    each a chain of ``nn.Linear`` (realnvp.py:26,28,41) whose default init is
    kaiming_uniform(a=sqrt(5)) for the weight then U(+-1/sqrt(fan_in)) for the
    bias.  Instantiating nn.Linear in that same order consumes the global torch
    RNG identically.
    """
    if seed is not None:
        torch.manual_seed(seed)
    params: Params = {}
    n_in = var_size + cond_size
    for i in range(n_layers):
        for net in ("t", "s"):
            sizes = [n_in] + list(hidden) + [var_size]
            for q, (wk, bk) in enumerate(net_keys(i, net, len(hidden))):
                lin = torch.nn.Linear(sizes[q], sizes[q + 1])
                params[wk] = lin.weight.detach().clone().to(dtype)
                params[bk] = lin.bias.detach().clone().to(dtype)
    return params


def param_order(n_layers: int, n_hidden: int) -> List[str]:
    """``nf.parameters()`` order: per layer t.0.w, t.0.b, ..., s.0.w, ...."""
    keys: List[str] = []
    for i in range(n_layers):
        for net in ("t", "s"):
            for wk, bk in net_keys(i, net, n_hidden):
                keys += [wk, bk]
    return keys


# --------------------------------------------------------------------------
# s/t conditioner and the coupling layer (reference realnvp.py:19-43, 73-129)
# --------------------------------------------------------------------------
def conditioner(u: torch.Tensor, params: Params, i: int, net: str, n_hidden: int,
                activation: str) -> torch.Tensor:
    """gen_network forward: Linear, act, ..., Linear -- realnvp.py:19-43."""
    h = u
    keys = net_keys(i, net, n_hidden)
    for q, (wk, bk) in enumerate(keys):
        h = torch.nn.functional.linear(h, params[wk], params[bk])
        if q < n_hidden:
            h = torch.tanh(h) if is_tanh(activation) else torch.relu(h)
    return h


def _net_input(X, C, mask):
    # realnvp.py:91-94 / 120-123: cat((X*mask, C)) or X*mask when C is None
    if C is not None:
        return torch.cat((X * mask[None, :], C), dim=1)
    return X * mask[None, :]


def coupling_f(X, C, params: Params, i: int, n_hidden: int, activation: str):
    """RealNVPLayer.f -- realnvp.py:73-101.  Returns (X_new [B,D], log_det [B])."""
    mask = layer_mask(X.shape[1], i)
    XC = _net_input(X, C, mask)
    T = conditioner(XC, params, i, "t", n_hidden, activation)
    S = conditioner(XC, params, i, "s", n_hidden, activation)
    X_new = (X * torch.exp(S) + T) * (1 - mask[None, :]) + X * mask[None, :]   # :99
    log_det = (S * (1 - mask[None, :])).sum(dim=-1)                             # :100
    return X_new, log_det


def coupling_g(X, C, params: Params, i: int, n_hidden: int, activation: str):
    """RealNVPLayer.g -- realnvp.py:104-129."""
    mask = layer_mask(X.shape[1], i)
    XC = _net_input(X, C, mask)
    T = conditioner(XC, params, i, "t", n_hidden, activation)
    S = conditioner(XC, params, i, "s", n_hidden, activation)
    return ((X - T) * torch.exp(-S)) * (1 - mask[None, :]) + X * mask[None, :]  # :128


# --------------------------------------------------------------------------
# the flow (reference nflow.py:90-145)
# --------------------------------------------------------------------------
def prior_log_prob(Z: torch.Tensor) -> torch.Tensor:
    """MultivariateNormal(0, I_D).log_prob(z) (prior built at realnvp.py:190-191)."""
    D = Z.shape[1]
    prior = torch.distributions.MultivariateNormal(torch.zeros(D, dtype=Z.dtype),
                                                   torch.eye(D, dtype=Z.dtype))
    return prior.log_prob(Z)


def flow_forward_rows(X, C, params: Params, n_layers: int, n_hidden: int, activation: str):
    """Body of NormalizingFlow.log_prob without the final mean -- nflow.py:107-115.

    Returns (z [B,D], logdet [B], logp [B]).
    """
    ll = None
    for i in range(n_layers):
        X, change = coupling_f(X, C, params, i, n_hidden, activation)
        ll = change if ll is None else ll + change
    logdet = ll
    logp = logdet + prior_log_prob(X)
    return X, logdet, logp


def flow_log_prob(X, C, params, n_layers, n_hidden, activation):
    """NormalizingFlow.log_prob: batch mean, 0-d tensor -- nflow.py:117."""
    return flow_forward_rows(X, C, params, n_layers, n_hidden, activation)[2].mean()


def flow_sample_from_noise(eps, C, params, n_layers, n_hidden, activation):
    """NormalizingFlow.sample with the prior draw replaced by ``eps`` -- nflow.py:141-143."""
    X = eps
    for i in reversed(range(n_layers)):
        X = coupling_g(X, C, params, i, n_hidden, activation)
    return X


def flow_sample(C, var_size, params, n_layers, n_hidden, activation):
    """NormalizingFlow.sample incl. the int path -- nflow.py:135-145.

    ``prior.sample((n,))`` is bit-identical to ``torch.randn(n, D)`` on the
    default generator, which is what is drawn here.
    """
    if type(C) == type(1):
        n, C = C, None
    else:
        n = len(C)
    eps = torch.randn(n, var_size)
    return flow_sample_from_noise(eps, C, params, n_layers, n_hidden, activation)


# --------------------------------------------------------------------------
# in-kernel prior draws of the product's sampling path (no reference line: upstream calls
# prior.sample((n,)) == torch.randn on the CPU generator, nflow.py:141).  The CUDA inverse kernels
# generate the same N(0,1) field from a counter-based generator keyed on the GLOBAL row index
# (probaforms_b200/csrc/rnvp_philox.cuh); this is its numpy restatement.
# --------------------------------------------------------------------------
def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11; Random123).  counter [..., 4], key [..., 2] uint32 arrays."""
    import numpy as np
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    for _ in range(10):
        p0 = c[0] * np.uint64(M0)
        p1 = c[2] * np.uint64(M1)
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c = [(hi1 ^ c[1] ^ k0) & np.uint64(MASK), lo1, (hi0 ^ c[3] ^ k1) & np.uint64(MASK), lo0]
        k0 = (k0 + np.uint64(W0)) & np.uint64(MASK)
        k1 = (k1 + np.uint64(W1)) & np.uint64(MASK)
    return np.stack(c, axis=-1).astype(np.uint32)


def philox_normal(seed: int, row_offset: int, n: int, var_size: int) -> torch.Tensor:
    """eps [n, var_size] fp32: element (r, j) from Philox(counter=(row_lo, row_hi, j//4, 0), key=seed) + Box-Muller,
    u = ((x >> 9) + 0.5) * 2^-23; outputs (0,1) -> features 4q, 4q+1 (cos, sin), (2,3) -> 4q+2, 4q+3."""
    import numpy as np
    nq = (var_size + 3) // 4
    rows = (np.arange(n, dtype=np.uint64) + np.uint64(row_offset))[:, None].repeat(nq, 1)
    ctr = np.zeros((n, nq, 4), dtype=np.uint64)
    ctr[..., 0] = rows & np.uint64(0xFFFFFFFF)
    ctr[..., 1] = rows >> np.uint64(32)
    ctr[..., 2] = np.arange(nq, dtype=np.uint64)[None, :]
    key = np.zeros((n, nq, 2), dtype=np.uint64)
    key[..., 0] = np.uint64(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint64((seed >> 32) & 0xFFFFFFFF)
    x = philox4x32_10(ctr, key)
    u = ((x >> np.uint32(9)).astype(np.float64) + 0.5) * 2.0 ** -23
    out = np.empty((n, nq, 4), dtype=np.float64)
    for a, b in ((0, 1), (2, 3)):
        r = np.sqrt(-2.0 * np.log(u[..., a]))
        out[..., a] = r * np.cos(2.0 * np.pi * u[..., b])
        out[..., b] = r * np.sin(2.0 * np.pi * u[..., b])
    return torch.from_numpy(out.reshape(n, nq * 4)[:, :var_size].astype(np.float32))


# --------------------------------------------------------------------------
# gradients and Adam (reference realnvp.py:205-207, 246-251)
# --------------------------------------------------------------------------
def loss_and_grads(X, C, params: Params, n_layers: int, n_hidden: int, activation: str):
    """loss = -log_prob (realnvp.py:246) and d loss / d theta for every tensor."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    loss = -flow_log_prob(X, C, leaves, n_layers, n_hidden, activation)
    order = param_order(n_layers, n_hidden)
    grads = torch.autograd.grad(loss, [leaves[k] for k in order], allow_unused=True)
    out = {}
    for k, g in zip(order, grads):
        out[k] = torch.zeros_like(params[k]) if g is None else g
    return loss.detach(), out


class AdamState:
    """State of torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay) -- realnvp.py:205-207."""

    def __init__(self, params: Params, lr: float, weight_decay: float = 0.0,
                 betas=(0.9, 0.999), eps: float = 1e-8):
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.step = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}


def adam_step(params: Params, grads: Params, st: AdamState) -> None:
    """One torch.optim.Adam step (single-tensor, non-amsgrad), in place.

    Restates torch/optim/adam.py ``_single_tensor_adam``: coupled L2
    (grad += wd*param), lerp for exp_avg, mul/addcmul for exp_avg_sq, python
    double bias corrections, ``param.addcdiv_(m, sqrt(v)/sqrt(bc2) + eps, -lr/bc1)``.
    """
    b1, b2 = st.betas
    st.step += 1
    bc1 = 1 - b1 ** st.step
    bc2 = 1 - b2 ** st.step
    step_size = st.lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for k, p in params.items():
        g = grads[k]
        if st.wd != 0:
            g = g.add(p, alpha=st.wd)
        st.m[k].lerp_(g, 1 - b1)
        st.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (st.v[k].sqrt() / bc2_sqrt).add_(st.eps)
        p.addcdiv_(st.m[k], denom, value=-step_size)


# --------------------------------------------------------------------------
# the fit loop (reference realnvp.py:210-262)
# --------------------------------------------------------------------------
def epoch_permutation(n: int) -> torch.Tensor:
    """Row order of one ``DataLoader(dataset, batch_size, shuffle=True)`` epoch.

    realnvp.py:237 builds a fresh DataLoader per epoch.  Iterating it draws two
    int64 from the global torch RNG -- the loader base seed
    (torch/utils/data/dataloader.py ``_BaseDataLoaderIter.__init__``) then the
    RandomSampler seed (sampler.py ``RandomSampler.__iter__``) -- and the batches
    are consecutive slices of ``randperm(n, generator=Generator(seed))``; the
    last partial batch is kept.
    """
    torch.empty((), dtype=torch.int64).random_()                 # loader base seed (unused)
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


def fit(X, C, params: Params, n_layers: int, n_hidden: int, activation: str,
        batch_size: int, n_epochs: int, lr: float, weight_decay: float = 0.0,
        state: Optional[AdamState] = None):
    """RealNVP.fit on already-initialised params -- realnvp.py:226-254.

    Returns (loss_history list of 0-d tensors, AdamState).  ``params`` is
    updated in place.
    """
    X = torch.as_tensor(X, dtype=torch.float32)
    C = None if C is None else torch.as_tensor(C, dtype=torch.float32)
    st = state or AdamState(params, lr, weight_decay)
    history = []
    n = X.shape[0]
    for _ in range(n_epochs):
        perm = epoch_permutation(n)
        for b0 in range(0, n, batch_size):
            idx = perm[b0:b0 + batch_size]
            Xb = X[idx]
            Cb = None if C is None else C[idx]
            loss, grads = loss_and_grads(Xb, Cb, params, n_layers, n_hidden, activation)
            adam_step(params, grads, st)
            history.append(loss)
    return history, st
