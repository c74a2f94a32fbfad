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
