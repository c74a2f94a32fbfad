"""Where does RealNVP.fit(X_host, C_host) spend its wall time?  (development aid)"""
import sys, time, torch
sys.path.insert(0, '/root/repo')
from probaforms_b200.models import RealNVP
import probaforms_b200.batching as B
D, Cd, L, H, bs, steps = 32, 8, 16, 128, 75776, 16
n = bs * steps
g = torch.Generator().manual_seed(7)
Xh = torch.randn(n, D, generator=g).pin_memory(); Ch = torch.randn(n, Cd, generator=g).pin_memory()
m = RealNVP(n_layers=L, hidden=(H,), batch_size=bs, n_epochs=1, lr=1e-4)
torch.manual_seed(0)
m.fit(Xh[:2 * bs], Ch[:2 * bs])
torch.cuda.synchronize()
T = {}
def tick(k, t0): T[k] = T.get(k, 0.0) + (time.perf_counter() - t0) * 1e3
# H2D alone
t = time.perf_counter(); Xd = Xh.to('cuda', non_blocking=True); Cd_ = Ch.to('cuda', non_blocking=True); torch.cuda.synchronize(); tick('h2d_194MB', t)
# shuffle alone
import ctypes as C
lib = m.nf._fused().lib
t = time.perf_counter(); sp = B.StreamingPermutation(lib, 123, n); tick('perm_create', t)
t = time.perf_counter(); sp.wait(bs); tick('perm_first_batch', t)
t = time.perf_counter(); sp.full(); tick('perm_rest', t)
# whole fit
t = time.perf_counter(); m.fit(Xh, Ch); torch.cuda.synchronize(); tick('fit_total', t)
# fit with wait instrumentation
orig_wait = B.StreamingPermutation.wait
def wait(self, upto):
    t0 = time.perf_counter(); r = orig_wait(self, upto); tick('wait_in_fit', t0); return r
B.StreamingPermutation.wait = wait
t = time.perf_counter(); m.fit(Xh, Ch); torch.cuda.synchronize(); tick('fit_total_2', t)
print({k: round(v, 2) for k, v in T.items()}, 'gpu-only estimate ms', steps * 1.585)
