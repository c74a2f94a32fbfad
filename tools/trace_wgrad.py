"""Wait accounting of CTA 0 of the tcgen05 weight-gradient sweep (rnvp_debug_set_trace): per role, the cycles spent in each
mbarrier wait and the role's total loop time.  Needs a library built with -DRNVP_WG_TRACE (make -C probaforms_b200/csrc
EXTRA=-DRNVP_WG_TRACE): the production build compiles the accounting away, because even the disabled accumulators cost the
kernel 15 % (register pressure)."""
import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo')
from probaforms_b200.models import RealNVPLayer, NormalizingFlow
dev = torch.device('cuda:0')
D, Cd, L, H, N = 32, 8, 16, 128, 75776
torch.manual_seed(0)
nf = NormalizingFlow([RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, (H,), 'tanh') for i in range(L)], None).to(dev)
eng = nf._fused()
X = torch.randn(N, D, device=dev); Cn = torch.randn(N, Cd, device=dev)
eng.zero_grads()
eng.backward(X, Cn, None, N, -1.0 / N)
torch.cuda.synchronize()
npad = (N + 255) // 256 * 256
ws = eng.workspace(N)
rec_ptr = C.c_void_p(ws.data_ptr() + 4 * npad * L * D)
buf = torch.zeros(4 * 2048 * 2, dtype=torch.int64, device=dev)
eng.lib.rnvp_debug_set_trace(C.c_void_p(buf.data_ptr()))
eng.lib.rnvp_wgrad_sweep(eng._desc, C.c_void_p(eng.packed.data_ptr()), npad, rec_ptr, C.c_void_p(eng.gpacked.data_ptr()), None)
torch.cuda.synchronize()
eng.lib.rnvp_debug_set_trace(None)
t = buf.cpu()[:40].view(5, 8).tolist()
nst = npad // 32 // (148 // (L * 2))
names = {0: ('issuer A', ['conv', 'accfree', 'afull', '-', '(dW1 issue)', '(dh issue)', '(dh issue+commit)']), 1: ('issuer B', ['-', 'accfree', 'afull']),
         2: ('converter', ['full', 'opfree']), 3: ('owner', ['full', 'dh', 'hfree', 'accfull']), 4: ('producer', ['empty'])}
print(f"stages per CTA ~{nst}; cycles per stage by role (total | waits):")
for r, (nm, ws_) in names.items():
    tot = t[r][7]
    print(f"  {nm:10s} total {tot / nst:7.0f} | " + "  ".join(f"{w} {t[r][k] / nst:6.0f}" for k, w in enumerate(ws_) if w != '-')
          + f"  | busy {(tot - sum(t[r][:4])) / nst:6.0f}")
