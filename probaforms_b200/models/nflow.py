"""Flow container with the reference's interface (probaforms/models/nflow.py:15-145).

``NormalizingFlow(layers, prior)`` keeps ``.layers`` (an ``nn.ModuleList`` of
``RealNVPLayer``) and ``.prior`` and the same three entry points --
``log_prob(X, C)`` -> 0-d batch mean, differentiable w.r.t. the parameters
(nflow.py:90-117); ``sample(C)`` (nflow.py:120-145) -- but instead of looping
``layer.f`` / ``layer.g`` in Python it sends the whole layer stack of a row
tile through ONE fused CUDA launch (probaforms_b200/csrc/rnvp_tile.cu).

Additive, opt-in extras (absent upstream): ``log_prob_rows``,
``forward_rows`` and ``sample_from_noise``.
"""
import weakref

import torch
import torch.nn as nn

from ..engine import FlowEngine


class InvertibleLayer(nn.Module):
    """Invertible map interface (reference nflow.py:15-67): ``f(X, C) -> (X_new, log_det)``,
    ``g(X, C) -> X_new``."""

    def __init__(self, var_size):
        super().__init__()
        self.var_size = var_size

    def f(self, X, C):
        raise NotImplementedError

    def g(self, X, C):
        raise NotImplementedError


class _LogProbMean(torch.autograd.Function):
    """mean_rows log p(x_row | c_row), with the parameter gradients produced by the fused
    forward+backward kernel during the forward call (activations are recomputed in-kernel, so
    nothing has to be kept for autograd)."""

    @staticmethod
    def forward(ctx, flow, X, C, *params):
        eng = flow._engine
        n = X.shape[0]
        ctx.spans = eng.tensor_spans
        ctx.shapes = [p.shape for p in params]
        if any(ctx.needs_input_grad[3:]):
            logp = torch.empty(n, dtype=torch.float32, device=eng.device)
            eng.zero_grads()
            eng.backward(X, C, None, n, 1.0 / n, logp_rows=logp)
            ctx.gflat = eng.unpack_grads()
            eng.zero_grads()                       # leave the accumulator clean for fit_step
        else:
            logp = eng.forward(X, C, want_z=False, want_logdet=False)[2]
            ctx.gflat = None
        return logp.mean()

    @staticmethod
    def backward(ctx, gout):
        g = ctx.gflat * gout
        grads = tuple(g[o:o + m].view(s) for (o, m), s in zip(ctx.spans, ctx.shapes))
        return (None, None, None) + grads


class NormalizingFlow(nn.Module):
    """Stack of coupling layers + prior (reference nflow.py:71-145)."""

    def __init__(self, layers, prior):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self.prior = prior
        self._engine = None
        self._link_layers()

    def _link_layers(self):
        # lets layer.f / layer.g route through the flow's fused engine as a 1-layer launch
        for i, layer in enumerate(self.layers):
            object.__setattr__(layer, "_flow_ref", (weakref.ref(self), i))

    # engines hold raw CUDA handles: never copy / pickle them, rebuild lazily instead
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        state.pop("_replicas", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._link_layers()

    # ------------------------------------------------------------------ fusion
    def _flow_shape(self):
        from .realnvp import RealNVPLayer
        if len(self.layers) == 0:
            raise RuntimeError("NormalizingFlow has no layers")
        first = self.layers[0]
        for i, layer in enumerate(self.layers):
            if not isinstance(layer, RealNVPLayer):
                raise NotImplementedError("only stacks of RealNVPLayer are fused by probaforms_b200")
            want = (torch.arange(layer.var_size) + i) % 2          # realnvp.py:199
            if layer.var_size != first.var_size or layer.cond_size != first.cond_size \
                    or layer.hidden != first.hidden or layer.activation != first.activation \
                    or not torch.equal(layer.mask.detach().cpu().to(want.dtype), want):
                raise NotImplementedError(
                    "the fused kernels need the RealNVP layout: identical layers with mask_i = (arange(D)+i)%2")
        return first.var_size, first.cond_size, len(self.layers), first.hidden, first.activation

    def _ordered_params(self):
        return [p for layer in self.layers for p in layer.parameters()]

    def _fused(self, repack=True):
        """The engine, with every nn.Parameter re-pointed at its span of the engine's flat buffer
        (so optimisers, state_dict and the kernels all see the same memory)."""
        params = self._ordered_params()
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("probaforms_b200 has no CPU path: move the flow to a CUDA device (sm_100a) first")
        eng = self._engine
        if eng is None or eng.device != dev:
            D, Cd, L, hidden, act = self._flow_shape()
            eng = FlowEngine(D, Cd, L, hidden, act, dev)
            self._engine = eng
        base = eng.flat.data_ptr()
        rebound = False
        for p, (off, numel) in zip(params, eng.tensor_spans):
            if p.data_ptr() != base + 4 * off or p.dtype != torch.float32:
                if p.numel() != numel:
                    raise RuntimeError("parameter shapes do not match the flow layout")
                view = eng.flat[off:off + numel].view(p.shape)
                view.copy_(p.data)
                p.data = view
                rebound = True
        if repack or rebound:
            eng.pack()
        return eng

    def _prep(self, X, C):
        eng = self._fused()
        X = torch.as_tensor(X, dtype=torch.float32, device=eng.device)
        if C is not None:
            C = torch.as_tensor(C, dtype=torch.float32, device=eng.device)
        return eng, X, C

    # ------------------------------------------------------------------ API
    def forward_rows(self, X, C=None):
        """(z [B,D], logdet [B], logp [B]) -- the loop body of nflow.py:107-115."""
        eng, X, C = self._prep(X, C)
        return eng.forward(X, C)

    def log_prob_rows(self, X, C=None):
        """Per-row log-density (no upstream equivalent; nflow.py:107-115 without the mean)."""
        eng, X, C = self._prep(X, C)
        return eng.forward(X, C, want_z=False, want_logdet=False)[2]

    def log_prob(self, X, C):
        """Batch-mean log-likelihood, 0-d tensor, differentiable w.r.t. parameters (nflow.py:90-117)."""
        eng, X, C = self._prep(X, C)
        if X.shape[0] == 0:
            return torch.full((), float("nan"), device=eng.device)
        return _LogProbMean.apply(self, X.contiguous(), None if C is None else C.contiguous(),
                                  *self._ordered_params())

    def sample_from_noise(self, eps, C=None):
        """Inverse pass on caller-supplied prior draws (parity mode of nflow.py:141-143)."""
        eng, eps, C = self._prep(eps, C)
        return eng.inverse(eps, C)

    def _replica(self, dev):
        """Engine holding a copy of this flow's weights on another CUDA device (multi-GPU sample / log-prob of one
        process: rows shard with no communication).  Refreshed from the primary at every call."""
        eng = self._fused(repack=False)
        dev = torch.device(dev)
        if dev == eng.device:
            return self._fused()
        reps = self.__dict__.setdefault("_replicas", {})
        rep = reps.get(dev)
        if rep is None:
            rep = FlowEngine(eng.D, eng.Cd, eng.L, eng.hidden, eng.activation, dev)
            reps[dev] = rep
        rep.flat.copy_(eng.flat)
        with torch.cuda.device(dev):
            rep.pack()
        return rep

    def sample(self, C, seed=None, row_offset=0, out=None):
        """nflow.py:120-145: ``C`` is a [n, cond_size] tensor or a Python int (unconditional).

        The prior draw (nflow.py:141) is generated inside the inverse kernel: latent element (r, j) is a function of
        (seed, row_offset + r, j) only, so row blocks produced on different GPUs or in different launches are the rows
        of one big request.  ``seed=None`` draws one int64 from torch's global CPU generator (reproducible under
        ``torch.manual_seed`` like upstream's ``prior.sample``).  ``sample_from_noise`` is the parity mode."""
        if type(C) == type(1):           # numpy ints deliberately do not qualify, as upstream (nflow.py:135)
            n, C = C, None
        else:
            n = len(C)
        eng = self._fused()
        if seed is None:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
        if C is not None:
            C = torch.as_tensor(C, dtype=torch.float32, device=eng.device)
        return eng.sample(n, C, seed=seed, row_offset=row_offset, out=out)

    def sample_many(self, C, n_draws, seed=None):
        """``n_draws`` independent ``sample(C)`` results as one [n_draws, n, D] tensor: the conditions are uploaded and
        validated once, every draw is one launch into its slice of the output (the notebooks' Monte-Carlo pattern
        ``for i in range(1000): model.sample(C)``, docs/examples/regression.ipynb cell 13).  Draw k consumes the global
        generator exactly like the k-th call of ``sample`` would."""
        if type(C) == type(1):
            n, C = C, None
        else:
            n = len(C)
        eng = self._fused()               # repack: in-place parameter edits (load_state_dict, external optimisers) count
        if C is not None:
            C = torch.as_tensor(C, dtype=torch.float32, device=eng.device)
        out = torch.empty(int(n_draws), n, eng.D, dtype=torch.float32, device=eng.device)
        for k in range(int(n_draws)):
            sd = int(torch.empty((), dtype=torch.int64).random_().item()) if seed is None else int(seed) + k
            eng.sample(n, C, seed=sd, out=out[k])
        return out
