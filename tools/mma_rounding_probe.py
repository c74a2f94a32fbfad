#!/usr/bin/env python
"""Probe the accumulation rounding of tcgen05.mma kind::tf32: with TF32-exact inputs every product is exact in
fp32, so the only error of a 1-pass run is the fp32 accumulation inside the tensor core."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probaforms_b200 import _lib  # noqa: E402


def to_tf32(x):
    i = x.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


lib = _lib.load()
dev = torch.device("cuda:0")
for K in (8, 16, 32, 64):
    N = 64
    g = torch.Generator().manual_seed(K)
    A = to_tf32(torch.rand(128, K, generator=g) + 0.5).to(dev)      # positive: sums grow monotonically
    B = to_tf32(torch.rand(N, K, generator=g) + 0.5).to(dev)
    D = torch.empty(128, N, device=dev)
    _lib.check(lib.rnvp_mma_selftest(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), N, K, 1, None), "selftest")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().T
    ulp = torch.ldexp(torch.ones_like(ref), torch.floor(torch.log2(ref)).int() - 23)
    e = (D.double() - ref) / ulp
    rn = (ref.float().double() - ref) / ulp
    print(f"K={K:3d} ({K // 8} accumulations): tensor-core error mean {float(e.mean()):+.3f} ulp, min {float(e.min()):+.2f}, max {float(e.max()):+.2f} | "
          f"single RN rounding of the exact sum: mean {float(rn.mean()):+.3f}, max |.| {float(rn.abs().max()):.2f}")
