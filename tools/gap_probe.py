"""What costs 0.17 ms between the streamed steps?  fit_step loop with (a) nothing, (b) a cross-stream event wait per step,
(c) an extra event record per step, (d) a continuous 12 MB H2D stream beside it.  (development aid)"""
import os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probaforms_b200.models import RealNVP

D, Cd, L, H, bs = 32, 8, 16, 128, 75776
rng = np.random.default_rng(0)
X = rng.standard_normal((4 * bs, D)); Cn = rng.standard_normal((4 * bs, Cd))
m = RealNVP(n_layers=L, hidden=(H,), batch_size=bs, n_epochs=1, lr=1e-4)
torch.manual_seed(0)
m.fit(X, Cn)
eng = m.nf._fused()
Xd = torch.randn(bs, D, device="cuda"); Cv = torch.randn(bs, Cd, device="cuda"); loss = torch.zeros(1, device="cuda")
side = torch.cuda.Stream()
hbig = torch.empty(12 << 18, pin_memory=True); dbig = torch.empty(12 << 18, device="cuda")


def run(kind, k=300):
    eng.zero_grads()
    for _ in range(400):                                    # reach the power-capped steady state first
        eng.fit_step(Xd, Cv, None, bs, bs, 1e-4, 0.0, loss)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    cur = torch.cuda.current_stream()
    for _ in range(k):
        if kind == "wait":
            ev = torch.cuda.Event()
            ev.record(side)
            cur.wait_event(ev)
        elif kind == "record":
            ev = torch.cuda.Event()
            ev.record(cur)
        elif kind == "timing-record":
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(cur)
        elif kind == "h2d-wait":
            with torch.cuda.stream(side):
                dbig.copy_(hbig, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(side)
            cur.wait_event(ev)                              # the step waits for ITS OWN upload (no prefetch)
        elif kind == "h2d-free":
            with torch.cuda.stream(side):
                dbig.copy_(hbig, non_blocking=True)         # an upload per step that nobody waits for
        eng.fit_step(Xd, Cv, None, bs, bs, 1e-4, 0.0, loss)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k


for kind in ("plain", "wait", "record", "timing-record", "h2d-free", "h2d-wait", "plain"):
    print("%-14s %.3f ms/step" % (kind, run(kind)), flush=True)
