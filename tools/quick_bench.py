#!/usr/bin/env python
"""Kernel-level timing of the three passes (log-prob / sample / fused fwd+bwd) for the BASELINE
shapes.  Development aid; bench.py is the contract benchmark."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, flops_per_row                       # noqa: E402
from probaforms_b200.models import RealNVPLayer, NormalizingFlow  # noqa: E402


def time_it(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="c2,c3,c4,c5")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--passes", default="fwd,inv,bwd")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak = 2 * 128 * sms * 1965e6 / 1e12
    for name in args.workloads.split(","):
        D, Cd, L, hidden, per_gpu, desc = WORKLOADS[name]
        n = args.rows or per_gpu * 4
        torch.manual_seed(0)
        nf = NormalizingFlow([RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, hidden, "tanh") for i in range(L)], None).to(dev)
        eng = nf._fused()
        X = torch.randn(n, D, device=dev)
        C = torch.randn(n, Cd, device=dev) if Cd else None
        lp = torch.empty(n, device=dev)
        out = torch.empty(n, D, device=dev)
        f_fwd, f_fit = flops_per_row(D, Cd, L, hidden[0])
        res = {"workload": name, "rows": n}
        for mode, key in ((0, "fwd"), (1, "inv"), (2, "bwd")):
            res[key + "_plan"] = eng.plan_info(mode)
        if "fwd" in args.passes:
            ms = time_it(lambda: eng.lib.rnvp_forward(eng._desc, eng.packed.data_ptr(), X.data_ptr(),
                                                      C.data_ptr() if Cd else None, None, n, 0, L, None, None,
                                                      lp.data_ptr(), None), args.reps)
            res["logprob_Mrows_s"] = n / ms / 1e3
            res["logprob_frac_fp32"] = n * f_fwd / (ms * 1e-3) / 1e12 / peak
        if "inv" in args.passes:
            ms = time_it(lambda: eng.sample(n, C, seed=1, out=out), args.reps)      # in-kernel prior draws (rnvp_sample)
            res["sample_Mrows_s"] = n / ms / 1e3
            res["sample_frac_fp32"] = n * f_fwd / (ms * 1e-3) / 1e12 / peak
        if "bwd" in args.passes:
            eng.zero_grads()
            ms = time_it(lambda: eng.backward(X, C, None, n, -1.0 / n), args.reps)
            res["fit_kernel_Mrows_s"] = n / ms / 1e3
            res["fit_frac_fp32"] = n * f_fit / (ms * 1e-3) / 1e12 / peak
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
