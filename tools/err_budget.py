#!/usr/bin/env python
"""Error budget of the GPU kernel families against an fp64 evaluation of the same flow
(err(ours, fp64) vs err(reference fp32, fp64)).  Uses the CPU oracle in fp32 and fp64."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import realnvp_oracle as O                            # noqa: E402
from probaforms_b200.models import RealNVPLayer, NormalizingFlow  # noqa: E402


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    dev = torch.device("cuda:0")
    for (D, Cd, L, H, seed) in [(32, 8, 16, 128, 10), (64, 16, 24, 128, 11), (64, 16, 24, 128, 3), (32, 8, 16, 128, 4)]:
        N = 2048
        params = O.init_params(D, Cd, L, (H,), seed=seed)
        p64 = {k: v.double() for k, v in params.items()}
        g = torch.Generator().manual_seed(seed + 1000)
        X = torch.randn(N, D, generator=g)
        C = torch.randn(N, Cd, generator=g)
        eps = torch.randn(N, D, generator=g)
        z64, ld64, lp64 = O.flow_forward_rows(X.double(), C.double(), p64, L, 1, "tanh")
        s64 = O.flow_sample_from_noise(eps.double(), C.double(), p64, L, 1, "tanh")
        z32, ld32, lp32 = O.flow_forward_rows(X, C, params, L, 1, "tanh")
        s32 = O.flow_sample_from_noise(eps, C, params, L, 1, "tanh")
        nf = NormalizingFlow([RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, (H,), "tanh") for i in range(L)], None)
        nf.load_state_dict(params)
        nf = nf.to(dev)
        eng = nf._fused()
        out = {"ref_fp32": (rel(z32, z64), rel(lp32, lp64), rel(s32, s64))}
        for name, path in (("tcgen05_tf32x3", 0), ("fp32_tile", 1)):
            eng.set_path(path)
            z, ld, lp = eng.forward(X.to(dev), C.to(dev))
            s = eng.inverse(eps.to(dev), C.to(dev))
            out[name] = (rel(z, z64), rel(lp, lp64), rel(s, s64))
        print(f"D={D} Cd={Cd} L={L} H={H} seed={seed} |logp|max={float(lp64.abs().max()):.0f}")
        for k, v in out.items():
            print(f"   {k:16s} z {v[0]:.2e}  logp {v[1]:.2e}  sample {v[2]:.2e}")


if __name__ == "__main__":
    main()
