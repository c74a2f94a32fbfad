"""RealNVP estimator with the reference's API (probaforms/models/realnvp.py:19-282).

Same constructor, ``fit(X, C)`` / ``sample(C)``, public attributes and
``state_dict`` key names (``nf.layers.{i}.nn_{t,s}.{2k}.{weight,bias}``) as
``probaforms.models.RealNVP``; the arithmetic runs in the fused sm_100a kernels
of ``librnvp_b200.so``.  Host code stays Python: epochs, batching (device-side,
consuming the torch RNG exactly like the reference's per-epoch
``DataLoader(shuffle=True)``), optimiser ownership.

Device: CUDA only.  ``cuda:{LOCAL_RANK}`` under torchrun, else the current CUDA
device; the reference's env var ``device`` (realnvp.py:12-15) is honoured when
it names a CUDA device and rejected otherwise -- there is no CPU fallback.
"""
import os
import weakref

import numpy as np
import torch
import torch.nn as nn

from .interfaces import GenModel
from .nflow import InvertibleLayer, NormalizingFlow
from ..engine import FlowEngine
from ..batching import PermutationPrefetcher, batch_bounds, epoch_seed, shard_bounds


def _default_device():
    env = os.environ.get("device")
    if env:
        dev = torch.device(env)
        if dev.type != "cuda":
            raise RuntimeError(f"probaforms_b200 is CUDA-only (sm_100a); env var device={env!r} is not supported")
        return dev
    if not torch.cuda.is_available():
        raise RuntimeError("probaforms_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")
    if "LOCAL_RANK" in os.environ:
        return torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    return torch.device("cuda", torch.cuda.current_device())


def gen_network(n_inputs, n_outputs, hidden=(10,), activation='tanh'):
    """Conditioner MLP: Linear, act, ..., Linear (reference realnvp.py:19-43).

    'tanh' selects Tanh; 'relu' and every other string select ReLU, as upstream (realnvp.py:32-37).
    The modules only hold the parameters (names, shapes, default init and RNG order identical to
    the reference); the fused kernels do the arithmetic.
    """
    widths = [n_inputs] + list(hidden)
    net = nn.Sequential()
    for a, b in zip(widths[:-1], widths[1:]):
        net.append(nn.Linear(a, b))
        net.append(nn.Tanh() if activation == 'tanh' else nn.ReLU())
    net.append(nn.Linear(widths[-1], n_outputs))
    return net


class RealNVPLayer(InvertibleLayer):
    """Affine coupling layer (reference realnvp.py:47-129).

    ``f``: y = (x*exp(s)+t)*(1-mask) + x*mask, log_det = sum(s*(1-mask))   (realnvp.py:99-100)
    ``g``: x = ((y-t)*exp(-s))*(1-mask) + y*mask                            (realnvp.py:128)
    with t, s = nn_t(u), nn_s(u), u = cat(x*mask, c).  Calls run as a one-layer launch of the fused
    kernels and are inference-only (no autograd through ``f`` / ``g``; training goes through
    ``NormalizingFlow.log_prob``).  Supported masks are the two RealNVP uses: (arange(D)+i)%2.
    """

    def __init__(self, var_size, cond_size, mask, hidden=(10,), activation='tanh'):
        super().__init__(var_size=var_size)
        self.cond_size = cond_size
        self.hidden = tuple(hidden)
        self.activation = activation
        self.mask = mask                           # plain attribute, not a buffer, as upstream (realnvp.py:68)
        self.nn_t = gen_network(var_size + cond_size, var_size, hidden, activation)   # nn_t first: fixes the
        self.nn_s = gen_network(var_size + cond_size, var_size, hidden, activation)   # init RNG order (:69-70)
        self._own_engine = None

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_own_engine"] = None
        state.pop("_flow_ref", None)
        return state

    def _parity(self):
        m = self.mask.detach().cpu().long()
        ar = torch.arange(self.var_size)
        if torch.equal(m, ar % 2):
            return 0
        if torch.equal(m, (ar + 1) % 2):
            return 1
        raise NotImplementedError("RealNVPLayer: only the alternating masks (arange(D)+i)%2 are supported")

    def _engine_and_range(self):
        ref = getattr(self, "_flow_ref", None)
        flow = ref[0]() if ref is not None else None
        if flow is not None:
            try:
                return flow._fused(), (ref[1], ref[1] + 1)
            except NotImplementedError:
                pass
        # stand-alone layer: a private two-layer descriptor (even mask, odd mask); this layer's
        # parameters are copied into the slot of its parity at every call
        par = self._parity()
        dev = next(self.parameters()).device
        eng = self._own_engine
        if eng is None or eng.device != dev:
            eng = FlowEngine(self.var_size, self.cond_size, 2, self.hidden, self.activation, dev)
            self._own_engine = eng
        n_t = len(eng.tensor_spans) // 2
        for p, (off, numel) in zip(self.parameters(), eng.tensor_spans[par * n_t:(par + 1) * n_t]):
            eng.flat[off:off + numel].copy_(p.detach().reshape(-1))
        eng.pack()
        return eng, (par, par + 1)

    def f(self, X, C=None):
        eng, rng = self._engine_and_range()
        X = torch.as_tensor(X, dtype=torch.float32, device=eng.device)
        C = None if C is None else torch.as_tensor(C, dtype=torch.float32, device=eng.device)
        y, log_det, _ = eng.forward(X, C, want_logp=False, layers=rng)
        return y, log_det

    def g(self, X, C=None):
        eng, rng = self._engine_and_range()
        X = torch.as_tensor(X, dtype=torch.float32, device=eng.device)
        C = None if C is None else torch.as_tensor(C, dtype=torch.float32, device=eng.device)
        return eng.inverse(X, C, layers=rng)


class FusedAdam(torch.optim.Adam):
    """``torch.optim.Adam`` whose ``step`` is one fused kernel over the flow's flat buffer.

    Constructed like the reference's optimiser (realnvp.py:205-207).  ``RealNVP.fit`` drives the
    engine directly (gradients never leave the packed accumulator); ``step()`` is for callers who
    run their own ``loss.backward()`` loop on ``nf.log_prob`` and reads the ``.grad`` tensors.
    Moments live in the engine (``exp_avg`` / ``exp_avg_sq`` flat buffers); ``state[p]`` exposes
    per-parameter views so ``state_dict()`` keeps working.
    """

    def __init__(self, flow, lr=1e-3, weight_decay=0.0):
        super().__init__(flow.parameters(), lr=lr, weight_decay=weight_decay)
        self._flow = weakref.ref(flow)

    def _publish_state(self, eng):
        for p, (off, numel) in zip(self._flow()._ordered_params(), eng.tensor_spans):
            st = self.state[p]
            st["step"] = torch.tensor(float(eng.adam_steps))
            st["exp_avg"] = eng.exp_avg[off:off + numel].view(p.shape)
            st["exp_avg_sq"] = eng.exp_avg_sq[off:off + numel].view(p.shape)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        flow = self._flow()
        eng = flow._fused(repack=False)
        group = self.param_groups[0]
        gflat = torch.zeros(eng.P, dtype=torch.float32, device=eng.device)
        for p, (off, numel) in zip(flow._ordered_params(), eng.tensor_spans):
            if p.grad is not None:
                gflat[off:off + numel].copy_(p.grad.reshape(-1))
        eng.adam_step(group["lr"], group["weight_decay"], betas=group["betas"], eps=group["eps"],
                      gflat_in=gflat, zero=False)
        self._publish_state(eng)
        return loss


class RealNVP(GenModel):
    """RealNVP normalizing flow (reference realnvp.py:133-282); parameters as upstream:

    n_layers, hidden, activation ('tanh' | 'relu'), batch_size, n_epochs, lr, weight_decay,
    verbose (>0: progress bar over epochs).
    """

    def __init__(self, n_layers=8, hidden=(10,), activation='tanh',
                 batch_size=32, n_epochs=10, lr=0.0001, weight_decay=0, verbose=0, shuffle='reference'):
        super().__init__()
        if shuffle not in ('reference', 'device'):
            raise ValueError("shuffle must be 'reference' or 'device'")
        # 'reference' (default): every epoch's batches have exactly the reference's composition (realnvp.py:237).
        # 'device' (additive, opt-in): the epoch order is a torch.randperm on the GPU, seeded from the same per-epoch
        # seed -- statistically the same training, not the same batches; it removes the sequential CPU shuffle
        # (13 ns per row) that bounds end-to-end throughput once several GPUs share one global batch.
        self.shuffle = shuffle
        self.n_layers = n_layers
        self.hidden = hidden
        self.activation = activation
        self.batch_size = batch_size
        self.n_epochs = n_epochs
        self.lr = lr
        self.weight_decay = weight_decay
        self.verbose = verbose

        self.prior = None
        self.nf = None
        self.opt = None

        self.loss_history = []
        self._device = None
        self._perm_host = [None, None]         # pinned staging buffers of the epoch row orders, reused across fits

    # ------------------------------------------------------------------ init
    def _model_init(self, X, C):
        """Lazy one-time construction (realnvp.py:180-207); a second ``fit`` warm-starts."""
        var_size = X.shape[1]
        cond_size = C.shape[1] if C is not None else 0
        if self._device is None:
            self._device = _default_device()
        dev = self._device
        if self.prior is None:
            self.prior = torch.distributions.MultivariateNormal(torch.zeros(var_size, device=dev),
                                                                torch.eye(var_size, device=dev))
        if self.nf is None:
            layers = [RealNVPLayer(var_size=var_size, cond_size=cond_size,
                                   mask=((torch.arange(var_size) + i) % 2),
                                   hidden=self.hidden, activation=self.activation)
                      for i in range(self.n_layers)]
            self.nf = NormalizingFlow(layers=layers, prior=self.prior).to(dev)
            for layer in self.nf.layers:
                layer.mask = layer.mask.to(dev)
            self.nf._fused()
            self.opt = FusedAdam(self.nf, lr=self.lr, weight_decay=self.weight_decay)

    @staticmethod
    def _to_device(A, dev):
        """numpy/torch -> contiguous float32 rows on the device (realnvp.py:226-228)."""
        if isinstance(A, torch.Tensor):
            # pinned host tensors upload asynchronously (the epoch's row order is computed meanwhile)
            nb = A.device.type == "cpu" and A.is_pinned()
            return A.to(device=dev, dtype=torch.float32, non_blocking=nb).contiguous()
        A = np.asarray(A)
        if A.dtype != np.float32:
            A = A.astype(np.float32)
        if A.ndim > 0:                                   # (ascontiguousarray would promote a 0-d value to 1-d)
            A = np.ascontiguousarray(A)
        return torch.from_numpy(A).to(dev)

    # ------------------------------------------------------------------ fit
    def fit(self, X, C=None):
        """Fit on X [n, var_size] (numpy), optional conditions C [n, cond_size] (realnvp.py:210-262).

        Per step: loss = -nf.log_prob(batch); zero_grad; backward; Adam step -- executed as one fused
        forward+backward launch plus one fused Adam launch.  ``loss_history`` gets one 0-d CPU tensor
        per step (as upstream) but is synchronised once per epoch instead of every step.

        Under an initialised ``torch.distributed`` group every rank passes the SAME X, C; each
        global batch of ``batch_size`` rows is split into contiguous per-rank slices and the packed
        gradient (+loss) buffer is all-reduced once per step (NCCL), which reproduces the
        single-process trajectory up to fp32 summation order.
        """
        self._model_init(X, C)
        dev = self._device
        dist = torch.distributed
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        n = X.shape[0]
        bs = int(self.batch_size)
        # identical row order on every rank (the sampler seed of rank 0 is broadcast); computed one epoch ahead on a
        # helper thread, the first one while the rows are uploaded
        eng = self.nf._fused()
        device_shuffle = getattr(self, "shuffle", "reference") == "device"
        perms = None if device_shuffle else PermutationPrefetcher(
            n, self.n_epochs, device=dev if world > 1 else None, lib=eng.lib, host_buffers=self._perm_host)
        Xd = self._to_device(X, dev)
        Cd = self._to_device(C, dev) if C is not None else None
        perm_dev = torch.empty(max(n, 1), dtype=torch.int64, device=dev)

        epochs = range(self.n_epochs)
        bar = None
        if self.verbose >= 1:
            from tqdm.auto import tqdm
            bar = tqdm(epochs, unit='epoch')
            epochs = bar
        eng.zero_grads()
        for _ in epochs:
            # the epoch's row order streams in from a helper thread (rnvp_perm_*, same order as the reference's
            # DataLoader): a step only waits for its own batch, the tail of the shuffle overlaps the GPU work
            bounds = batch_bounds(n, bs)
            losses = torch.empty(len(bounds), dtype=torch.float32, device=dev)
            loss_ptr = losses.data_ptr()
            if device_shuffle:
                gen = torch.Generator(device=dev)
                gen.manual_seed(epoch_seed(device=dev if world > 1 else None) & 0x7FFFFFFFFFFFFFFF)
                perm_dev = torch.randperm(n, device=dev, generator=gen)
                stream, copied = None, n
            else:
                stream, copied = perms.next_stream(), 0
            perm_ptr = perm_dev.data_ptr()
            for s, (b0, nb) in enumerate(bounds):           # last partial batch is kept (drop_last=False)
                if copied < b0 + nb:                        # upload whatever is final by now, at least this batch
                    upto = max(b0 + nb, min(n, stream.available()))
                    host = stream.wait(upto)
                    perm_dev[copied:upto].copy_(host[copied:upto], non_blocking=True)
                    copied = upto
                lo, hi = shard_bounds(b0, nb, rank, world)
                # raw device addresses instead of tensor slices: the host side of a 32-row step is the bottleneck
                eng.fit_step(Xd, Cd, perm_ptr + 8 * lo, hi - lo, nb, self.lr, self.weight_decay,
                             loss_ptr + 4 * s, world=world)
            host = losses.cpu()                             # the epoch's only device->host sync
            self.loss_history.extend(host.unbind(0))
            if bar is not None:
                bar.set_description(f"loss: {float(host[-1]):.4f}")
        self.opt._publish_state(eng)

    # ------------------------------------------------------------------ sample
    def sample(self, C=100, n_draws=None):
        """Draw rows for the given conditions [n, cond_size], or ``C`` rows if it is a Python int
        (realnvp.py:265-282).  Returns a float32 numpy array [n, var_size].

        ``n_draws=k`` (additive, not upstream) returns [k, n, var_size]: k independent draws for the same conditions
        with one upload of ``C`` and one download of the result -- the notebooks' ``for i in range(1000):
        model.sample(C)`` loop as a single call."""
        if type(C) != type(1):
            C = self._to_device(C, self._device)
        if n_draws is not None:
            return self.nf.sample_many(C, n_draws).cpu().detach().numpy()
        X = self.nf.sample(C).cpu().detach().numpy()
        return X
