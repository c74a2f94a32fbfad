import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo')
from probaforms_b200.models import RealNVPLayer, NormalizingFlow
from probaforms_b200 import _lib
dev = torch.device('cuda:0')
D, Cd, L, H = 32, 8, 16, 128
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K1P = (D // 2 + Cd + 7) // 8 * 8
torch.manual_seed(0)
nf = NormalizingFlow([RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, (H,), 'tanh') for i in range(L)], None).to(dev)
eng = nf._fused()
g = torch.Generator(device=dev).manual_seed(1)
h = torch.randn(L, N, 2, H, device=dev, generator=g)
d1 = torch.randn(L, N, 2, H, device=dev, generator=g)
d2 = torch.randn(L, N, 2, D // 2, device=dev, generator=g)
u = torch.randn(L, N, K1P, device=dev, generator=g)
u[:, :, D // 2 + Cd:] = 0
P = lambda t: C.c_void_p(t.data_ptr())
rec = eng.lib.rnvp_wgrad_record_floats(eng._desc)
R = torch.zeros(L, N, rec, device=dev)
R[:, :, :2 * H] = d1.reshape(L, N, 2 * H)
R[:, :, 2 * H:4 * H] = h.reshape(L, N, 2 * H)
R[:, :, 4 * H:4 * H + K1P] = u
R[:, :, 4 * H + K1P:4 * H + K1P + D] = d2.reshape(L, N, D)
print('record floats', rec)
# blocked layout of the C ABI: [L][N/32][rec/4][32 slots][4], slot = row ^ 4*(group & 1)
Rb = R.view(L, N // 32, 32, rec // 4, 4).permute(0, 1, 3, 2, 4).contiguous()
perm = torch.arange(32, device=dev) ^ 4
Rb[:, :, 1::2] = Rb[:, :, 1::2][:, :, :, perm]
R = Rb
def run():
    _lib.check(eng.lib.rnvp_wgrad_sweep(eng._desc, N, P(R), P(eng.gpacked), None), 'wgrad')
eng.zero_grads(); run(); torch.cuda.synchronize()
gflat = eng.unpack_grads()
spans = eng.tensor_spans
names = [n for n, _ in nf.named_parameters()]
err = 0.0; mx = 0.0
for i in range(L):
    par = i & 1
    for net, nm in enumerate('ts'):
        dW1 = d1[i, :, net].double().T @ u[i, :, :D // 2 + Cd].double()
        gW1 = gflat[spans[names.index(f'layers.{i}.nn_{nm}.0.weight')][0]:][:H * (D + Cd)].view(H, D + Cd).double()
        cols = list(range(1 - par, D, 2)) + list(range(D, D + Cd))
        err = max(err, float((gW1[:, cols] - dW1).abs().max())); mx = max(mx, float(dW1.abs().max()))
        dW2 = d2[i, :, net].double().T @ h[i, :, net].double()
        gW2 = gflat[spans[names.index(f'layers.{i}.nn_{nm}.2.weight')][0]:][:D * H].view(D, H).double()
        err = max(err, float((gW2[par::2] - dW2).abs().max())); mx = max(mx, float(dW2.abs().max()))
        gb1 = gflat[spans[names.index(f'layers.{i}.nn_{nm}.0.bias')][0]:][:H].double()
        err = max(err, float((gb1 - d1[i, :, net].double().sum(0)).abs().max()))
        gb2 = gflat[spans[names.index(f'layers.{i}.nn_{nm}.2.bias')][0]:][:D].double()
        err = max(err, float((gb2[par::2] - d2[i, :, net].double().sum(0)).abs().max()))
print('max abs err', err, 'max abs val', mx, 'rel', err / mx)
eng.zero_grads()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
run(); torch.cuda.synchronize()
a.record()
for _ in range(5): run()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print('wgrad sweep ms', ms, 'for N =', N, ' -> GB/s', R.numel() * 4 / ms / 1e6)
