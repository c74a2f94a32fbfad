import sys, torch
sys.path.insert(0, '/root/repo')
from probaforms_b200.models import RealNVPLayer, NormalizingFlow
dev = torch.device('cuda:0')
D, Cd, L, H = 32, 8, 16, 128
N = 65536
out = {}
for path in (1, 0):
    torch.manual_seed(0)
    nf = NormalizingFlow([RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, (H,), 'tanh') for i in range(L)], None).to(dev)
    eng = nf._fused()
    eng.set_path(path)
    gen = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(4 * N, D, device=dev, generator=gen); C = torch.randn(4 * N, Cd, device=dev, generator=gen)
    perm = torch.randint(0, 4 * N, (40 * N,), device=dev, generator=gen)
    losses = torch.zeros(40, device=dev)
    eng.zero_grads()
    for s in range(40):
        eng.fit_step(X, C, perm[s * N:(s + 1) * N], N, N, 1e-4, 0.0, losses[s:s + 1])
    out[path] = losses.cpu()
print('fp32  ', [round(float(v), 3) for v in out[1][::4]])
print('hybrid', [round(float(v), 3) for v in out[0][::4]])
