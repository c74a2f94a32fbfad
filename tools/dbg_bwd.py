"""Per-tensor gradient comparison: tcgen05 fit path (path 0) against the FP32 tile kernels (path 1)."""
import sys, torch
sys.path.insert(0, '/root/repo')
from probaforms_b200.models import RealNVPLayer, NormalizingFlow
dev = torch.device('cuda:0')
D, Cd, L, H, N = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (32, 8, 2, 64, 512))]
res = {}
for path in (1, 0):
    torch.manual_seed(0)
    nf = NormalizingFlow([RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, (H,), 'tanh') for i in range(L)], None).to(dev)
    eng = nf._fused()
    eng.set_path(path)
    g = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(N, D, device=dev, generator=g)
    C = torch.randn(N, Cd, device=dev, generator=g) if Cd else None
    eng.zero_grads()
    eng.backward(X, C, None, N, -1.0 / N)
    torch.cuda.synchronize()
    res[path] = (eng.unpack_grads().cpu().double(), float(eng.loss_slot))
    names = [n for n, _ in nf.named_parameters()]
    spans = eng.tensor_spans
print('loss', res[1][1], res[0][1])
worst = 0.0
for k, nm in enumerate(names):
    o, n = spans[k]
    a, b = res[1][0][o:o + n], res[0][0][o:o + n]
    r = float((a - b).abs().max() / a.abs().max().clamp_min(1e-30))
    worst = max(worst, r)
    if r > 2e-5 or '-v' in sys.argv:
        print(f'{nm:32s} rel {r:.3e}  |ref| {float(a.abs().max()):.3e} |new| {float(b.abs().max()):.3e}')
print('worst rel', worst)
