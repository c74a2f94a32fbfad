"""Where does RealNVP.fit(X_numpy, C_numpy) spend its wall time?  (development aid)  Host gather rate by thread count, the
pieces of one streamed step, and the whole call with both shuffles."""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from probaforms_b200.models import RealNVP
from probaforms_b200 import _lib
import probaforms_b200.ingest as I
lib = _lib.load()
D, Cd, L, H, bs = 32, 8, 16, 128, 75776
n = bs * 131
rng = np.random.default_rng(0)
blk = rng.standard_normal((1 << 20, D + Cd))
XC = np.tile(blk, ((n + len(blk) - 1) // len(blk), 1))[:n]
X, Cn = np.ascontiguousarray(XC[:, :D]), np.ascontiguousarray(XC[:, D:])
idx = rng.permutation(n).astype(np.int64)
hx = torch.empty(bs, D, pin_memory=True); hc = torch.empty(bs, Cd, pin_memory=True)
for thr in (1, 4, 8, 12, 15):
    ts = []
    for k in range(12):
        t0 = time.perf_counter()
        lib.rnvp_host_gather_xc(C.c_void_p(X.ctypes.data), 1, D, C.c_void_p(Cn.ctypes.data), 1, Cd, C.c_void_p(idx[k * bs:].ctypes.data), 0, bs,
                                C.c_void_p(hx.data_ptr()), C.c_void_p(hc.data_ptr()), thr)
        ts.append(time.perf_counter() - t0)
    print(f"gather f64 {bs} rows, {thr:2d} threads: median {sorted(ts)[6] * 1e3:.3f} ms  min {min(ts) * 1e3:.3f} ms")
dx = torch.empty(bs, D, device='cuda'); dc = torch.empty(bs, Cd, device='cuda')
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    dx.copy_(hx, non_blocking=True); dc.copy_(hc, non_blocking=True)
torch.cuda.synchronize(); print(f"H2D of one step's rows: {(time.perf_counter() - t0) * 100:.3f} ms")
print("host_threads()", I.host_threads())
for shuffle in ("reference", "device"):
    for ingest in ("auto", "stream", "resident"):
        m = RealNVP(n_layers=L, hidden=(H,), batch_size=bs, n_epochs=1, lr=1e-4, shuffle=shuffle, ingest=ingest)
        torch.manual_seed(0)
        m.fit(X[:4 * bs], Cn[:4 * bs])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m.fit(X, Cn)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"fit {n} rows shuffle={shuffle:9s} ingest={ingest:8s}: {dt * 1e3:7.1f} ms = {n / dt / 1e6:.1f} M rows/s  ({dt / 131 * 1e3:.3f} ms/step)")
