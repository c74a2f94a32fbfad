import sys, ctypes as C, torch
sys.path.insert(0, '/root/repo')
from probaforms_b200 import _lib
lib = _lib.load(); dev = torch.device('cuda:0')
for N, K in [(16, 16), (16, 8), (64, 32), (32, 16)]:
    g = torch.Generator().manual_seed(1)
    A = torch.randn(128, K, generator=g).to(dev); B = torch.randn(N, K, generator=g).to(dev)
    ref = A.double() @ B.double().T
    for passes in (1, 4, 5):
        D = torch.full((128, N), float('nan'), device=dev)
        _lib.check(lib.rnvp_mma_selftest(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), N, K, passes, None), 'st')
        torch.cuda.synchronize()
        print(N, K, passes, float((D.double() - ref).abs().max() / ref.abs().max()), float(D.abs().max()))
