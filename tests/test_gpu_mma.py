"""tcgen05 primitive self-test on the GPU: TMEM staging, smem descriptors, instruction descriptor,
commit/mbarrier and the TF32x3 split accuracy (fp32-grade) are checked against an fp64 product."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(64, 32), (16, 32), (256, 32), (32, 56), (128, 8), (16, 64)])
def test_tcgen05_selftest(N, K):
    from probaforms_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N * 100 + K)
    A = torch.randn(128, K, generator=g).to(dev)
    B = torch.randn(N, K, generator=g).to(dev)
    ref = (A.double() @ B.double().T)
    scale = float(ref.abs().max())
    errs = {}
    for passes in (1, 3):
        D = torch.full((128, N), float("nan"), device=dev)
        rc = lib.rnvp_mma_selftest(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()),
                                   N, K, passes, None)
        _lib.check(rc, "rnvp_mma_selftest")
        torch.cuda.synchronize()
        errs[passes] = float((D.double() - ref).abs().max()) / scale
    fp32 = float(((A @ B.T).double() - ref).abs().max()) / scale
    print(f"N={N} K={K}: tf32 {errs[1]:.2e}  tf32x3 {errs[3]:.2e}  torch-fp32 {fp32:.2e}")
    assert errs[1] < 2e-3           # plain TF32: ~2^-11 per operand
    assert errs[3] < 2e-6           # split: fp32-grade


@pytest.mark.parametrize("shape", [(32, 8, 16, 128), (64, 16, 24, 128), (32, 0, 3, 64), (32, 5, 4, 256), (64, 16, 2, 32),
                                   (128, 32, 8, 512), (128, 0, 2, 64), (128, 7, 3, 96), (64, 16, 3, 512),
                                   # padded shapes: D != 2 * DH (odd D included), Cd up to 16 on the D <= 32 kernels
                                   (16, 4, 3, 64), (24, 8, 3, 96), (48, 16, 2, 128), (9, 0, 3, 32), (100, 30, 2, 128), (32, 16, 2, 64)])
@pytest.mark.parametrize("N", [1, 255, 257, 40000])
def test_tcgen05_flow_matches_fp32_kernels(shape, N):
    """The tcgen05 (TF32x3) forward / inverse kernels against the FP32-FMA tile kernels of the same
    library and the g(f(x)) round trip; tolerance rel 1e-5 like every other parity test."""
    from probaforms_b200.models import RealNVPLayer, NormalizingFlow
    D, Cd, L, H = shape
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    nf = NormalizingFlow([RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, (H,), "tanh") for i in range(L)], None).to(dev)
    eng = nf._fused()
    assert eng.plan_info(0)["kernel_family"] == 2
    g = torch.Generator().manual_seed(N)
    X = torch.randn(N, D, generator=g).to(dev)
    Cn = torch.randn(N, Cd, generator=g).to(dev) if Cd else None
    eps = torch.randn(N, D, generator=g).to(dev)
    z, ld, lp = eng.forward(X, Cn)
    x_back = eng.inverse(z, Cn)
    xs = eng.inverse(eps, Cn)
    eng.set_path(1)
    assert eng.plan_info(0)["kernel_family"] == 0
    z0, ld0, lp0 = eng.forward(X, Cn)
    xs0 = eng.inverse(eps, Cn)
    eng.set_path(0)

    def rel(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
    print(f"shape {shape} N={N}: z {rel(z, z0):.2e} logdet {rel(ld, ld0):.2e} logp {rel(lp, lp0):.2e} sample {rel(xs, xs0):.2e}")
    assert rel(z, z0) < 1e-5 and rel(lp, lp0) < 1e-5 and rel(ld, ld0) < 1e-5
    assert rel(xs, xs0) < 1e-5
    assert float((x_back - X).abs().max()) < 1e-4 * max(1.0, float(X.abs().max()))


def test_fit_steps_on_tcgen05_path_track_fp32_path():
    """Regression: the Adam kernel must refresh the TF32 hi/lo weight images of the tcgen05 kernels every step
    (a stale image trains nothing).  20 fused fit steps on both kernel families give the same loss curve."""
    from probaforms_b200.models import RealNVPLayer, NormalizingFlow
    dev = torch.device("cuda:0")
    D, Cd, L, H, N = 32, 8, 6, 64, 4096
    curves = {}
    for path in (1, 0):
        torch.manual_seed(0)
        nf = NormalizingFlow([RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, (H,), "tanh") for i in range(L)], None).to(dev)
        eng = nf._fused()
        eng.set_path(path)
        assert eng.plan_info(2)["kernel_family"] == (2 if path == 0 else 0)
        g = torch.Generator(device=dev).manual_seed(1)
        X = torch.randn(4 * N, D, device=dev, generator=g)
        C = torch.randn(4 * N, Cd, device=dev, generator=g)
        perm = torch.randint(0, 4 * N, (20 * N,), device=dev, generator=g)
        losses = torch.zeros(20, device=dev)
        eng.zero_grads()
        for s in range(20):
            eng.fit_step(X, C, perm[s * N:(s + 1) * N], N, N, 1e-3, 0.0, losses[s:s + 1])
        curves[path] = losses.cpu()
    assert float(curves[1][-1]) < float(curves[1][0]) - 1.0          # it does train
    assert torch.allclose(curves[0], curves[1], rtol=1e-4, atol=1e-4), (curves[0], curves[1])


@pytest.mark.parametrize("D,Cd,H,N", [(32, 8, 64, 2048), (32, 8, 128, 4096 + 32), (32, 8, 96, 960), (32, 8, 32, 64),
                                      (64, 16, 128, 2048), (128, 32, 128, 1024), (128, 32, 512, 4096)])
def test_wgrad_sweep_through_the_abi_matches_fp64(D, Cd, H, N):
    """rnvp_wgrad_sweep alone (C ABI): blocked records [L][N/32][rec/4][32][4] built on the host side of the ABI, gradients
    against an fp64 evaluation of dW1 = delta1^T u, dW2 = delta2^T h with delta1 = (delta2 W2) * (1 - h^2).  H = 64 puts
    both nets into one 128-lane block (block-diagonal W2 image), H = 128 gives one block per net, H = 48 / 16 partial blocks."""
    import ctypes as C
    from probaforms_b200 import _lib
    from probaforms_b200.models import RealNVPLayer, NormalizingFlow
    dev = torch.device("cuda:0")
    L = 3
    K1P = (D // 2 + Cd + 7) // 8 * 8
    torch.manual_seed(0)
    nf = NormalizingFlow([RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, (H,), "tanh") for i in range(L)], None).to(dev)
    eng = nf._fused()
    rec = eng.lib.rnvp_wgrad_record_floats(eng._desc)
    assert rec == 2 * H + K1P + D
    g = torch.Generator(device=dev).manual_seed(1)
    h = torch.tanh(torch.randn(L, N, 2, H, device=dev, generator=g))
    d2 = torch.randn(L, N, 2, D // 2, device=dev, generator=g)
    u = torch.randn(L, N, K1P, device=dev, generator=g)
    u[:, :, D // 2 + Cd:] = 0
    R = torch.cat([h.reshape(L, N, 2 * H), u, d2.reshape(L, N, D)], dim=2)
    Rb = R.view(L, N // 32, 32, rec // 4, 4).permute(0, 1, 3, 2, 4).contiguous()
    for q in range(1, 8):                                                            # slot = row ^ (group & 7)
        Rb[:, :, q::8] = Rb[:, :, q::8][:, :, :, torch.arange(32, device=dev) ^ q]
    eng.zero_grads()
    P = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(eng.lib.rnvp_wgrad_sweep(eng._desc, P(eng.packed), N, P(Rb), P(eng.gpacked), None), "rnvp_wgrad_sweep")
    torch.cuda.synchronize()
    gflat = eng.unpack_grads().double()
    names = [n for n, _ in nf.named_parameters()]
    spans = dict(zip(names, eng.tensor_spans))
    sd = {n: p.detach().double() for n, p in nf.named_parameters()}
    worst = 0.0
    for i in range(L):
        par = i & 1
        xk_cols = list(range(1 - par, D, 2)) + list(range(D, D + Cd))
        for net, nm in enumerate("ts"):
            W2 = sd[f"layers.{i}.nn_{nm}.2.weight"][par::2]                        # rows of the transformed features
            d1 = (d2[i, :, net].double() @ W2) * (1 - h[i, :, net].double() ** 2)
            want = {f"layers.{i}.nn_{nm}.0.weight": (d1.T @ u[i, :, :D // 2 + Cd].double(), None, xk_cols),
                    f"layers.{i}.nn_{nm}.0.bias": (d1.sum(0), None, None),
                    f"layers.{i}.nn_{nm}.2.weight": (d2[i, :, net].double().T @ h[i, :, net].double(), slice(par, None, 2), None),
                    f"layers.{i}.nn_{nm}.2.bias": (d2[i, :, net].double().sum(0), slice(par, None, 2), None)}
            for name, (ref, rows, cols) in want.items():
                off, numel = spans[name]
                got = gflat[off:off + numel].view(sd[name].shape)
                if rows is not None:
                    got = got[rows]
                if cols is not None:
                    got = got[:, cols]
                worst = max(worst, float((got - ref).abs().max() / ref.abs().max()))
    print("wgrad sweep vs fp64: worst rel", worst)
    assert worst < 5e-6
    # argument checks
    assert eng.lib.rnvp_wgrad_sweep(eng._desc, P(eng.packed), 33, P(Rb), P(eng.gpacked), None) < 0
    assert eng.lib.rnvp_wgrad_sweep(eng._desc, None, N, P(Rb), P(eng.gpacked), None) < 0
