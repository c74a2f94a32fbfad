"""Does the c3 fit step slow down under sustained load (power / clocks)?  (development aid)"""
import os, sys, time, subprocess, threading
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
samples, stop = [], False


def smi():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True).stdout.strip()
        samples.append((time.perf_counter(), out))
        time.sleep(0.05)


th = threading.Thread(target=smi); th.start()
time.sleep(1.0)
eng.zero_grads()
t_start = time.perf_counter()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(31)]
evs[0].record()
for w in range(30):
    for _ in range(100):
        eng.fit_step(Xd, Cv, None, bs, bs, 1e-4, 0.0, loss)
    evs[w + 1].record()
    evs[w + 1].synchronize()
torch.cuda.synchronize()
t_end = time.perf_counter()
stop = True; th.join()
print("ms/step per 100-step window:", " ".join("%.3f" % (evs[i].elapsed_time(evs[i + 1]) / 100) for i in range(30)))
for t, s in samples:
    if t_start - 0.3 < t < t_end + 0.2:
        print("  t=%6.2f s  sm_mhz,power_w,temp,reasons = %s" % (t - t_start, s))
