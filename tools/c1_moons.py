"""configs[0]: README make_moons RealNVP(lr=0.01, n_epochs=100) fit + sample, 1000 rows (reference CPU: ~43 s fit)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from probaforms_b200.models import RealNVP
from sklearn.datasets import make_moons
X, y = make_moons(n_samples=1000, noise=0.1, random_state=0)
C = y.reshape(-1, 1)
torch.manual_seed(0)
m = RealNVP(lr=0.01, n_epochs=100)
t = time.perf_counter(); m.fit(X, C); torch.cuda.synchronize(); dt = time.perf_counter() - t
t = time.perf_counter(); S = m.sample(C); ds = time.perf_counter() - t
print({"fit_s": round(dt, 3), "steps": len(m.loss_history), "rows_per_s": round(100000 / dt), "us_per_step": round(dt / len(m.loss_history) * 1e6, 1),
       "final_loss": float(m.loss_history[-1]), "sample_ms": round(ds * 1e3, 3), "sample_shape": S.shape})
t = time.perf_counter(); m.fit(X, C); torch.cuda.synchronize(); dt = time.perf_counter() - t
print({"second_fit_s": round(dt, 3), "final_loss": float(m.loss_history[-1])})
