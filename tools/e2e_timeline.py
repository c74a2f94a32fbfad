"""Timeline of the streamed RealNVP.fit(X_numpy, C_numpy) (development aid): where each step's wall time goes.

Prints (1) pinned H2D / D2H rates by transfer size, alone and with the host gather running beside them, (2) per-step
means of the StepStreamer worker phases (slot wait, order wait, gather, enqueue), of the fit loop's wait in next(), and
of the GPU-side upload / kernel durations taken from CUDA events, (3) the whole-call rate."""
import os
import sys
import threading
import time
import ctypes as C

os.environ["RNVP_INGEST_TRACE"] = "1"
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probaforms_b200.models import RealNVP
from probaforms_b200 import _lib
import probaforms_b200.ingest as I

lib = _lib.load()
D, Cd, L, H, bs = 32, 8, 16, 128, 75776
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n = bs * steps
rng = np.random.default_rng(0)
blk = rng.standard_normal((1 << 20, D + Cd))
XC = np.tile(blk, ((n + len(blk) - 1) // len(blk), 1))[:n]
X, Cn = np.ascontiguousarray(XC[:, :D]), np.ascontiguousarray(XC[:, D:])
del XC
print("cpu_count", os.cpu_count(), "host_threads", I.host_threads(), flush=True)

# ---- (1) bus rates
for mb in (2, 12, 64, 256):
    h = torch.empty(mb << 18, pin_memory=True)
    d = torch.empty(mb << 18, device="cuda")
    for name, dst, src in (("H2D", d, h), ("D2H", h, d)):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"{name} {mb:4d} MB pinned: {dt * 1e3:7.3f} ms = {mb / 1024 / dt * 1.048576:.2f} GB/s")
h = torch.empty(64 << 18, pin_memory=True)
d = torch.empty(64 << 18, device="cuda")
h2 = torch.empty(64 << 18, pin_memory=True)
d2 = torch.empty(64 << 18, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"H2D + D2H 64 MB each, concurrently: {dt * 1e3:.3f} ms = {2 * 64 / 1024 / dt * 1.048576:.2f} GB/s total")

# H2D while 15 threads gather
idx = rng.permutation(n).astype(np.int64)
hx = torch.empty(bs, D, pin_memory=True)
hc = torch.empty(bs, Cd, pin_memory=True)
stop = False


def gather_loop():
    k = 0
    while not stop:
        lib.rnvp_host_gather_xc(C.c_void_p(X.ctypes.data), 1, D, C.c_void_p(Cn.ctypes.data), 1, Cd,
                                C.c_void_p(idx[(k % (steps - 1)) * bs:].ctypes.data), 0, bs,
                                C.c_void_p(hx.data_ptr()), C.c_void_p(hc.data_ptr()), I.host_threads())
        k += 1


ts = []
for k in range(20):
    t0 = time.perf_counter()
    lib.rnvp_host_gather_xc(C.c_void_p(X.ctypes.data), 1, D, C.c_void_p(Cn.ctypes.data), 1, Cd, C.c_void_p(idx[k * bs:].ctypes.data), 0, bs,
                            C.c_void_p(hx.data_ptr()), C.c_void_p(hc.data_ptr()), I.host_threads())
    ts.append(time.perf_counter() - t0)
print(f"gather f64 {bs} rows alone: median {sorted(ts)[10] * 1e3:.3f} ms")
th = threading.Thread(target=gather_loop)
th.start()
hbig = torch.empty(12 << 18, pin_memory=True)
dbig = torch.empty(12 << 18, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    dbig.copy_(hbig, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
stop = True
th.join()
print(f"H2D 12 MB while the gather pool runs: {dt * 1e3:.3f} ms = {12 / 1024 / dt * 1.048576:.2f} GB/s")

# ---- (2) instrumented streamed fit
orig_next, orig_release = I.StepStreamer.next, I.StepStreamer.release
log = {"next_wait": [], "ev": [], "streamer": None}


def next_(self):
    t0 = time.perf_counter()
    r = orig_next(self)
    log["next_wait"].append(time.perf_counter() - t0)
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    log["ev"].append([e0, None])
    log["streamer"] = self
    return r


def release_(self, slot):
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    log["ev"][-1][1] = e1
    orig_release(self, slot)


I.StepStreamer.next, I.StepStreamer.release = next_, release_

for shuffle in ("reference",):
    m = RealNVP(n_layers=L, hidden=(H,), batch_size=bs, n_epochs=1, lr=1e-4, shuffle=shuffle, ingest="stream")
    torch.manual_seed(0)
    m.fit(X[:4 * bs], Cn[:4 * bs])
    log["next_wait"].clear(), log["ev"].clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.fit(X, Cn)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tr = np.array([t[:4] for t in log["streamer"].trace])
    starts = np.array([t[4] for t in log["streamer"].trace])
    kern = np.array([a.elapsed_time(b) for a, b in log["ev"] if b is not None])
    gaps = np.array([log["ev"][i][1].elapsed_time(log["ev"][i + 1][0]) for i in range(len(log["ev"]) - 1)])
    print(f"\nstream fit shuffle={shuffle}: {dt * 1e3:.1f} ms for {steps} steps = {n / dt / 1e6:.1f} M rows/s ({dt / steps * 1e3:.3f} ms/step)")
    print("  worker per step [ms]: slot wait %.3f  order wait %.3f  gather %.3f  enqueue %.3f   (first step starts %.1f ms after fit())"
          % (*(tr[2:].mean(0) * 1e3), (starts[0] - t0) * 1e3))
    print("  worker step period [ms]: mean %.3f  p50 %.3f  p90 %.3f" % (np.diff(starts).mean() * 1e3, np.median(np.diff(starts)) * 1e3,
                                                                      np.percentile(np.diff(starts), 90) * 1e3))
    print("  fit loop: next() wait mean %.3f ms;  GPU: kernels of a step %.3f ms, idle gap between steps %.3f ms"
          % (np.mean(log["next_wait"][2:]) * 1e3, kern[2:].mean(), gaps[2:].mean()))
    print("  first 6 steps worker phases [ms]:", np.round(tr[:6] * 1e3, 2).tolist())
    te = log["streamer"].trace_events
    base = log["ev"][40][0]
    print("  GPU timeline of steps 40..47 [ms after kernels(40) start]: upload start, upload end | kernels start, kernels end")
    for k in range(40, 48):
        print("    step %d: upload %.3f .. %.3f | kernels %.3f .. %.3f" % (k, base.elapsed_time(te[k][0]), base.elapsed_time(te[k][1]),
                                                                      base.elapsed_time(log["ev"][k][0]), base.elapsed_time(log["ev"][k][1])))

I.StepStreamer.next, I.StepStreamer.release = orig_next, orig_release
# ---- (4) sample() end to end
Cs = np.ascontiguousarray(Cn[:1 << 20], dtype=np.float32)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    Xs = m.sample(Cs)
    dt = time.perf_counter() - t0
    print(f"sample 1,048,576 rows -> numpy: {dt * 1e3:.2f} ms = {len(Cs) / dt / 1e6:.1f} M rows/s  (pinned result: {type(Xs.base).__name__})")
    del Xs
