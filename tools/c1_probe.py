import sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from probaforms_b200.models import RealNVP
from sklearn.datasets import make_moons
Xm, ym = make_moons(n_samples=1000, noise=0.1, random_state=0)
big = torch.empty(10 << 30, dtype=torch.uint8, device='cuda'); del big
torch.cuda.empty_cache()
warm = RealNVP(lr=0.01, n_epochs=2); warm.fit(Xm, ym.reshape(-1, 1)); warm.sample(ym.reshape(-1, 1)); torch.cuda.synchronize()
for rep in range(5):
    torch.manual_seed(0)
    t0 = time.perf_counter()
    mm = RealNVP(lr=0.01, n_epochs=100)
    t1 = time.perf_counter()
    mm._model_init(Xm, ym.reshape(-1, 1)); eng = mm.nf._fused(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    mm.fit(Xm, ym.reshape(-1, 1)); torch.cuda.synchronize()
    t3 = time.perf_counter()
    print("ctor %.1f ms, init+engine %.1f ms, fit %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
