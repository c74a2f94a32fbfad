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


def run_pipelined(k=300, pieces=1):
    """Double-buffered uploads the way a streamer does them: the upload of step j+1 is released when step j-1 has finished,
    so it runs beside the kernels of step j."""
    eng.zero_grads()
    for _ in range(400):
        eng.fit_step(Xd, Cv, None, bs, bs, 1e-4, 0.0, loss)
    torch.cuda.synchronize()
    cur = torch.cuda.current_stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ups = []
    a.record()
    fin = [None, None]                                       # kernels of the step that last used buffer j & 1
    upl = None
    n = dbig.numel() // pieces
    for j in range(k):
        with torch.cuda.stream(side):
            if fin[j & 1] is not None:
                side.wait_event(fin[j & 1])
            u0 = torch.cuda.Event(enable_timing=True); u0.record(side)
            for q in range(pieces):
                dbig[q * n:(q + 1) * n].copy_(hbig[q * n:(q + 1) * n], non_blocking=True)
            u1 = torch.cuda.Event(enable_timing=True); u1.record(side)
            ups.append((u0, u1))
        if upl is not None:
            cur.wait_event(upl[1])                           # step j consumes the upload enqueued one iteration earlier
        eng.fit_step(Xd, Cv, None, bs, bs, 1e-4, 0.0, loss)
        f = torch.cuda.Event(); f.record(cur)
        fin[j & 1] = f
        upl = ups[-1]
    b.record()
    torch.cuda.synchronize()
    dur = np.array([u0.elapsed_time(u1) for u0, u1 in ups[20:]])
    return a.elapsed_time(b) / k, dur.mean()


for kind in ("plain", "h2d-wait"):
    print("%-14s %.3f ms/step" % (kind, run(kind)), flush=True)
for pieces in (1, 8):
    t, d = run_pipelined(pieces=pieces)
    print("pipelined uploads beside the kernels (%d piece(s)): %.3f ms/step, upload takes %.3f ms" % (pieces, t, d), flush=True)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ctypes as C
from probaforms_b200 import _lib
import probaforms_b200.ingest as I
lib = _lib.load()
nbig = bs * 40
XX = np.tile(rng.standard_normal((1 << 18, D)), (nbig // (1 << 18) + 1, 1))[:nbig]
CC = np.tile(rng.standard_normal((1 << 18, Cd)), (nbig // (1 << 18) + 1, 1))[:nbig]
idx = rng.permutation(nbig).astype(np.int64)
hx = torch.empty(bs, D, pin_memory=True); hc = torch.empty(bs, Cd, pin_memory=True)
stop = False


def gather_loop():
    k = 0
    while not stop:
        lib.rnvp_host_gather_xc(C.c_void_p(XX.ctypes.data), 1, D, C.c_void_p(CC.ctypes.data), 1, Cd,
                                C.c_void_p(idx[(k % 39) * bs:].ctypes.data), 0, bs,
                                C.c_void_p(hx.data_ptr()), C.c_void_p(hc.data_ptr()), I.host_threads())
        k += 1
        time.sleep(0.0008)


th = threading.Thread(target=gather_loop); th.start()
t, d = run_pipelined()
print("pipelined uploads + host gather pool busy half of the time: %.3f ms/step, upload takes %.3f ms" % (t, d), flush=True)
stop = True; th.join()


def run_one_thread_streamer(k=300, slots=3):
    """Gather -> upload -> kernels from ONE host thread: the gather of step j runs while the GPU works on steps j-1, j-2."""
    eng.zero_grads()
    for _ in range(400):
        eng.fit_step(Xd, Cv, None, bs, bs, 1e-4, 0.0, loss)
    torch.cuda.synchronize()
    cur = torch.cuda.current_stream()
    hxs = [torch.empty(bs, D, pin_memory=True) for _ in range(slots)]
    hcs = [torch.empty(bs, Cd, pin_memory=True) for _ in range(slots)]
    dxs = [torch.empty(bs, D, device="cuda") for _ in range(slots)]
    dcs = [torch.empty(bs, Cd, device="cuda") for _ in range(slots)]
    fin = [None] * slots
    ups = []
    thr = I.host_threads()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tg = 0.0
    for j in range(k):
        s = j % slots
        if fin[s] is not None:
            fin[s].synchronize()
        t0 = time.perf_counter()
        lib.rnvp_host_gather_xc(C.c_void_p(XX.ctypes.data), 1, D, C.c_void_p(CC.ctypes.data), 1, Cd,
                                C.c_void_p(idx[(j % 39) * bs:].ctypes.data), 0, bs,
                                C.c_void_p(hxs[s].data_ptr()), C.c_void_p(hcs[s].data_ptr()), thr)
        tg += time.perf_counter() - t0
        with torch.cuda.stream(side):
            u0 = torch.cuda.Event(enable_timing=True); u0.record(side)
            dxs[s].copy_(hxs[s], non_blocking=True)
            dcs[s].copy_(hcs[s], non_blocking=True)
            u1 = torch.cuda.Event(enable_timing=True); u1.record(side)
            ups.append((u0, u1))
        cur.wait_event(u1)
        eng.fit_step(dxs[s], dcs[s], None, bs, bs, 1e-4, 0.0, loss)
        f = torch.cuda.Event(); f.record(cur)
        fin[s] = f
    b.record()
    torch.cuda.synchronize()
    dur = np.array([u0.elapsed_time(u1) for u0, u1 in ups[20:]])
    return a.elapsed_time(b) / k, dur.mean(), tg / k * 1e3


for slots in (2, 3, 4):
    t, d, g = run_one_thread_streamer(slots=slots)
    print("one-thread streamer, %d slots: %.3f ms/step, upload takes %.3f ms, gather %.3f ms" % (slots, t, d, g), flush=True)
