"""README c1 fit: wall time through the API and GPU time of a 32-row step (development aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probaforms_b200.models import RealNVP
from sklearn.datasets import make_moons
Xm, ym = make_moons(n_samples=1000, noise=0.1, random_state=0)
w = RealNVP(lr=0.01, n_epochs=2); w.fit(Xm, ym.reshape(-1, 1)); w.sample(ym.reshape(-1, 1).astype(np.float32))
for rep in range(3):
    torch.manual_seed(0)
    c1 = RealNVP(lr=0.01, n_epochs=100)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c1.fit(Xm, ym.reshape(-1, 1))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    hist = torch.stack(c1.loss_history)
    print("c1 fit: %.3f s, %.1f us/step, last-epoch mean loss %.4f" % (dt, dt / len(hist) * 1e6, float(hist[-32:].mean())))
e1 = c1.nf._fused()
Xs = torch.tensor(Xm, dtype=torch.float32, device="cuda"); Cs1 = torch.tensor(ym.reshape(-1, 1), dtype=torch.float32, device="cuda")
l1 = torch.zeros(32, device="cuda")
perm = torch.randperm(1000, device="cuda")
e1.zero_grads()
for _ in range(3):
    e1.fit_epoch(Xs, Cs1, perm, 1000, 32, 0.01, 0.0, l1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    e1.fit_epoch(Xs, Cs1, perm, 1000, 32, 0.01, 0.0, l1)
b.record(); torch.cuda.synchronize()
print("c1 fit_epoch (fused steps) GPU-timed: %.2f us/step" % (a.elapsed_time(b) / 320 * 1e3))
a.record()
for _ in range(200):
    e1.fit_step(Xs, Cs1, perm[:32], 32, 32, 0.01, 0.0, l1)
b.record(); torch.cuda.synchronize()
print("c1 fit_step (two launches) GPU-timed: %.2f us/step" % (a.elapsed_time(b) / 200 * 1e3))
