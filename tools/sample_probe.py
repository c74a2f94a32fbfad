"""Pieces of RealNVP.sample(C_host) -> numpy on the c3 flow, the host enqueue cost of one fit step, and the README c1 fit
(development aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probaforms_b200.models import RealNVP
import probaforms_b200.ingest as I

D, Cd, L, H, bs = 32, 8, 16, 128, 75776
rng = np.random.default_rng(0)
X = rng.standard_normal((4 * bs, D)); Cn = rng.standard_normal((4 * bs, Cd))
m = RealNVP(n_layers=L, hidden=(H,), batch_size=bs, n_epochs=1, lr=1e-4)
torch.manual_seed(0)
m.fit(X, Cn)
eng = m.nf._fused()
n = 1 << 20
Cs = np.ascontiguousarray(rng.standard_normal((n, Cd)), dtype=np.float32)


def wall(f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


Cdv = m._to_device(Cs, m._device)
print("upload C (1M x 8 f32): %.3f ms" % wall(lambda: m._to_device(Cs, m._device)))
out = torch.empty(n, D, device="cuda")
print("kernel 1M rows: %.3f ms" % wall(lambda: eng.sample(n, Cdv, seed=1, out=out)))
arr, flat = I.RESULTS.lend((n, D))
print("D2H 134 MB into lent pinned: %.3f ms" % wall(lambda: flat.view(n, D).copy_(out, non_blocking=True)))
del arr, flat
for ch in (32768, 65536, 131072, 262144, 524288, 1048576):
    def f():
        r = m._sample_to_host(eng, Cs, 0, n, [1], False, chunk_rows=ch)
        del r
    print("sample_to_host chunk %7d: %.3f ms" % (ch, wall(f)))
def g():
    r = m.sample(Cs); del r
print("RealNVP.sample: %.3f ms" % wall(g))

# host enqueue cost of a fit step
Xd = torch.randn(bs, D, device="cuda"); Cv = torch.randn(bs, Cd, device="cuda"); loss = torch.zeros(1, device="cuda")
eng.zero_grads()
for _ in range(5):
    eng.fit_step(Xd, Cv, None, bs, bs, 1e-4, 0.0, loss)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    eng.fit_step(Xd, Cv, None, bs, bs, 1e-4, 0.0, loss)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("fit_step: host enqueue %.3f ms/step, total %.3f ms/step" % ((t1 - t0) / 20 * 1e3, (t2 - t0) / 20 * 1e3))

# c1
from sklearn.datasets import make_moons
Xm, ym = make_moons(n_samples=1000, noise=0.1, random_state=0)
w = RealNVP(lr=0.01, n_epochs=2); w.fit(Xm, ym.reshape(-1, 1)); w.sample(ym.reshape(-1, 1).astype(np.float32))
for rep in range(2):
    torch.manual_seed(0)
    c1 = RealNVP(lr=0.01, n_epochs=100)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c1.fit(Xm, ym.reshape(-1, 1))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    hist = torch.stack(c1.loss_history)
    print("c1 fit: %.3f s, %.1f us/step, last-epoch mean loss %.4f" % (dt, dt / len(hist) * 1e6, float(hist[-32:].mean())))
e1 = c1.nf._fused()
Xs = torch.tensor(Xm[:32], dtype=torch.float32, device="cuda"); Cs1 = torch.tensor(ym[:32].reshape(-1, 1), dtype=torch.float32, device="cuda")
l1 = torch.zeros(1, device="cuda")
e1.zero_grads()
for _ in range(20):
    e1.fit_step(Xs, Cs1, None, 32, 32, 0.01, 0.0, l1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200):
    e1.fit_step(Xs, Cs1, None, 32, 32, 0.01, 0.0, l1)
b.record(); torch.cuda.synchronize()
print("c1 fit_step (32 rows) GPU-timed loop: %.2f us/step" % (a.elapsed_time(b) / 200 * 1e3))
