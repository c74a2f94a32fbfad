"""Timeline of CTA 0's backward sweep in the tcgen05 fit kernel (rnvp_debug_set_trace)."""
import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo')
from probaforms_b200.models import RealNVPLayer, NormalizingFlow
dev = torch.device('cuda:0')
D, Cd, L, H, N = 32, 8, 16, 128, 65536
torch.manual_seed(0)
nf = NormalizingFlow([RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, (H,), 'tanh') for i in range(L)], None).to(dev)
eng = nf._fused()
X = torch.randn(N, D, device=dev); Cn = torch.randn(N, Cd, device=dev)
eng.zero_grads()
eng.backward(X, Cn, None, N, -1.0 / N)
buf = torch.zeros(4 * 2048 * 2, dtype=torch.int64, device=dev)
eng.lib.rnvp_debug_set_trace(C.c_void_p(buf.data_ptr()))
eng.backward(X, Cn, None, N, -1.0 / N)
torch.cuda.synchronize()
eng.lib.rnvp_debug_set_trace(None)
ev = buf.cpu().view(4, 2048, 2)
t0 = min(int(ev[w, 0, 1]) for w in range(4) if ev[w, 0, 1] > 0)
rows = []
for w in range(4):
    for k in range(2048):
        tag, t = int(ev[w, k, 0]), int(ev[w, k, 1])
        if t == 0: break
        rows.append((t - t0, w, tag))
rows.sort()
# print one middle layer (the 5th traced layer of the first pair)
starts = [t for t, w, tag in rows if w == 0 and 100 <= tag < 200]
lo, hi = starts[4], starts[6]
prev = {}
for t, w, tag in rows:
    if lo <= t < hi:
        print(f"{t - lo:7d}  {'  ' * 9 * w}{['T0', 'T1', 'MMA0', 'MMA1'][w]} {tag:4d}  (+{t - prev.get(w, t)})")
    prev[w] = t
print('layer period (clks):', [b - a for a, b in zip(starts[:-1], starts[1:])][:16])
