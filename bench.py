#!/usr/bin/env python
"""bench.py -- RealNVP rows/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N ...            # reference CPU path (oracle port)

Workload (default ``c3``, BASELINE.json configs[2], the configuration the metric's "fit ... at
1/2/4/8 B200" is quoted on): RealNVP fit on synthetic N(0,1) rows, D=32, Cd=8, 16 coupling
layers, hidden (128,), tanh, fp32.  A *step* is one optimisation step of the hot path over one
batch: tensor-core forward+backward sweep launch, weight-gradient sweep launch, gradient all-reduce
(N>1), fused Adam launch (reference realnvp.py:246-251).  Weak scaling: 75,776 rows per GPU per step
(= 2 row-tile pairs for each of the 148 persistent CTAs; --rows-per-gpu overrides).

value      rows/s, whole job, inputs resident in HBM (a different random batch of a resident
           data set larger than L2 every step -- no L2 flush needed).
e2e        same metric through the public API ``RealNVP.fit(X, C)`` with pinned HOST arrays:
           H2D of every step's rows and D2H of the losses inside the timed region.
roofline   dominant kernel (fused fwd+bwd): algorithmic mask-aware GEMM flops per launch
           (909,312 / row for c3) / its mean duration from CUDA events inside the
           timed region, against the FP32-FMA peak 2*128*SMs*max clock (the binding pipe; the
           path is compute-bound, HBM fraction is reported beside it).
cpu_baseline  the oracle port (same ATen ops as the reference) timed on the host cores on a
           bounded sample, rank 0, N=1 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (D, Cd, L, hidden, rows per GPU per step, description)
    "c2": (2, 1, 8, (10,), 1 << 20, "configs[1]: 2-D moons flow, 1-D condition, L=8, H=10"),
    # 75,776 = 148 SMs x 2 pairs of 128-row tiles x 256 rows: every persistent CTA of the tcgen05 fit kernel gets exactly two
    # row-tile pairs per step (no tail wave; at 65,536 rows 108 of the 148 CTAs get 2 pairs, the other 40 one)
    "c3": (32, 8, 16, (128,), 75776, "configs[2]: fit, 32-D rows, 8-D condition, L=16, H=128"),
    "c4": (64, 16, 24, (128,), 32768, "configs[3]: 64-D rows, 16-D condition, L=24, H=128 (H assumed)"),
    "c5": (128, 32, 8, (512,), 16384, "configs[4]: 128-D rows, 32-D condition, L=8 (assumed), H=512"),
}


def flops_per_row(D, Cd, L, H):
    """Mask-aware GEMM flops per row, 2 per MAC."""
    fwd = bwd = 0
    for i in range(L):
        nT = (D - (i & 1) + 1) // 2
        nK = D - nT
        fwd += 4 * H * (nK + Cd + nT)
        bwd += 2 * (2 * H * 2 * nT + 2 * H * (nK + Cd) + (2 * H * nK if i > 0 else 0))
    return fwd, fwd + bwd


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_step_rate(D, Cd, L, hidden, rows, reps, warm):
    """The oracle port (ATen CPU ops, all host threads): full optimisation steps -> rows/s."""
    import torch
    from oracle import realnvp_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    params = O.init_params(D, Cd, L, hidden, seed=0)
    st = O.AdamState(params, lr=1e-4)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(rows, D, generator=g)
    C = torch.randn(rows, Cd, generator=g) if Cd else None
    times = []
    for it in range(warm + reps):
        t0 = time.perf_counter()
        _, grads = O.loss_and_grads(X, C, params, L, len(hidden), "tanh")
        O.adam_step(params, grads, st)
        dt = time.perf_counter() - t0
        if it >= warm:
            times.append(dt)
    return rows * len(times) / sum(times), torch.get_num_threads(), sum(times) / len(times)


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference is
    pure Python/torch and cannot travel to the GPU box; the port issues the same ATen calls)."""
    D, Cd, L, hidden, per_gpu, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = 16384
    rate, cores, step_s = cpu_port_step_rate(D, Cd, L, hidden, rows, reps=args.steps, warm=args.warmup)
    line = {
        "impl": "reference", "metric": "RealNVP fit rows/sec (fwd+bwd+Adam)", "value": rate, "unit": "rows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + " -- " + desc, "rows_per_step": rows,
                   "note": "reference CPU path = oracle port (same ATen ops), bounded sample"},
        "cpu_baseline": {"value": rate, "unit": "rows/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} optimisation steps of {rows} rows"},
        "e2e": {"value": rate, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--rows-per-gpu", type=int, default=0, help="rows per GPU per step (default per workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the short c1 / c4 / c5 measurements")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    from probaforms_b200.models import RealNVP, RealNVPLayer, NormalizingFlow

    D, Cd, L, hidden, per_gpu, desc = wl
    if args.rows_per_gpu:
        per_gpu = args.rows_per_gpu
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to fd 1 when the communicator is created
        # (NCCL_DEBUG=VERSION/WARN in the environment), so fd 1 points at stderr until the first collective is done
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            import datetime
            # a rank that dies or skips a collective must fail the run in minutes, not hang it
            dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    K, W = args.steps, max(args.warmup, 3)

    # ---- flow with random-init weights of the named architecture (identical on every rank)
    torch.manual_seed(0)
    layers = [RealNVPLayer(D, Cd, (torch.arange(D) + i) % 2, hidden, "tanh") for i in range(L)]
    nf = NormalizingFlow(layers, prior=None).to(dev)
    eng = nf._fused()

    # ---- resident synthetic data set, larger than L2 (126 MB): random rows gathered every step
    n_res = max(4 * per_gpu, (768 << 20) // (4 * (D + Cd)))
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    X = torch.randn(n_res, D, device=dev, generator=gen)
    C = torch.randn(n_res, Cd, device=dev, generator=gen) if Cd else None
    perm = torch.randint(0, n_res, ((K + W) * per_gpu,), device=dev, generator=gen)
    losses = torch.zeros(K + W, device=dev)
    n_global = per_gpu * world
    lr, wd = 1e-4, 0.0

    def step(s):
        eng.fit_step(X, C, perm[s * per_gpu:(s + 1) * per_gpu], per_gpu, n_global, lr, wd,
                     losses[s:s + 1], world=world)

    eng.zero_grads()
    for s in range(W):
        step(s)
    torch.cuda.synchronize()

    # ---- timed region: exactly K steps, barrier + sync on both sides, CUDA events, max over ranks
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = eng.launches
    ev0.record()
    for s in range(K):
        a, b = kev[s]
        a.record()
        eng.backward(X, C, perm[(W + s) * per_gpu:(W + s + 1) * per_gpu], per_gpu, -1.0 / n_global)
        b.record()
        if world > 1:
            dist.all_reduce(eng._gbuf)
        eng.adam_step(lr, wd, loss_dst=losses[W + s:W + s + 1], loss_scale=-1.0 / n_global)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launches - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    kern_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in kev) / K], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(ms)

    # ---- tcgen05 fit path: rnvp_backward = rnvp_mma_kernel<..,2> (forward + backward sweeps) + rnvp_wgrad_kernel.
    # Time the weight-gradient sweep alone on the records the last step left in the workspace (it accumulates into the
    # gradient buffer, which is re-zeroed afterwards); the tcgen05 kernel's share is the difference.
    wgrad_ms = None
    if eng.fit_on_tensor_cores:
        import ctypes as ct
        npad = (per_gpu + 255) // 256 * 256
        ws = eng.workspace(per_gpu)
        rec_off = npad * L * D                         # floats of the forward stash that precedes the records
        rec_ptr = ct.c_void_p(ws.data_ptr() + 4 * rec_off)
        wa, wb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rc = 0
        wa.record()
        for _ in range(K):
            rc |= eng.lib.rnvp_wgrad_sweep(eng._desc, ct.c_void_p(eng.packed.data_ptr()), npad, rec_ptr,
                                           ct.c_void_p(eng.gpacked.data_ptr()), None)
        wb.record()
        torch.cuda.synchronize()
        if rc == 0:
            wgrad_ms = wa.elapsed_time(wb) / K
        eng.zero_grads()
    rows_per_s = n_global * K / (total_ms * 1e-3)
    final_loss = float(losses[W + K - 1])

    # ---- the other two passes of the metric on the same flow: per-row log-prob and sample (no communication)
    def time_pass(fn, rows, reps=5):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return rows * world / (float(t) * 1e-3), float(t)

    def pass_rates(engine, Xr, Cr):
        n = Xr.shape[0]
        lp = torch.empty(n, device=dev)
        out = torch.empty_like(Xr)
        r_lp, ms_lp = time_pass(lambda: engine.lib.rnvp_forward(
            engine._desc, engine.packed.data_ptr(), Xr.data_ptr(), Cr.data_ptr() if Cr is not None else None, None, n,
            0, engine.L, None, None, lp.data_ptr(), None), n)
        r_s, ms_s = time_pass(lambda: engine.inverse(Xr, Cr, out=out), n)
        return r_lp, ms_lp, r_s, ms_s

    n_pass = min(n_res, 1 << 20)
    lp_rate, lp_ms, s_rate, s_ms = pass_rates(eng, X[:n_pass], None if C is None else C[:n_pass])
    fam = {0: "fp32 tile kernel", 1: "small-flow kernel", 2: "tcgen05 TF32x3 kernel"}[eng.plan_info(0)["kernel_family"]]

    # ---- configs[1] (c2): 2-D moons flow, per-row log-prob and sample, 16.7 M rows per GPU
    D2, Cd2, L2, hid2, _, desc2 = WORKLOADS["c2"]
    torch.manual_seed(0)
    nf2 = NormalizingFlow([RealNVPLayer(D2, Cd2, (torch.arange(D2) + i) % 2, hid2, "tanh") for i in range(L2)],
                          prior=None).to(dev)
    eng2 = nf2._fused()
    n2 = 1 << 24
    X2 = torch.randn(n2, D2, device=dev, generator=gen)
    C2 = (torch.rand(n2, Cd2, device=dev, generator=gen) > 0.5).float()
    c2_lp, c2_lp_ms, c2_s, c2_s_ms = pass_rates(eng2, X2, C2)
    del X2, C2

    # ---- the other named configurations, one short measurement each (rank-local rows, no communication):
    # configs[3] (c4) per-row log-density, configs[4] (c5) wide fit step + log-density, configs[0] (c1) README moons fit
    others = {}
    if args.workload == "c3" and not args.no_others:
        for name in ("c4", "c5"):
            Do, Cdo, Lo, hido, per_o, desco = WORKLOADS[name]
            torch.manual_seed(0)
            nfo = NormalizingFlow([RealNVPLayer(Do, Cdo, (torch.arange(Do) + i) % 2, hido, "tanh") for i in range(Lo)],
                                  prior=None).to(dev)
            engo = nfo._fused()
            n_o = per_o * (8 if name == "c4" else 2)
            Xo = torch.randn(n_o, Do, device=dev, generator=gen)
            Co = torch.randn(n_o, Cdo, device=dev, generator=gen)
            o_lp, o_lp_ms, o_s, o_s_ms = pass_rates(engo, Xo, Co)
            fo_fwd, fo_fit = flops_per_row(Do, Cdo, Lo, hido[0])
            fam_o = {0: "fp32 tile kernel", 1: "small-flow kernel", 2: "tcgen05 TF32x3 kernel"}[engo.plan_info(0)["kernel_family"]]
            ent = {"log_prob_rows_s": o_lp, "sample_rows_s": o_s, "rows_per_launch": n_o, "kernel": fam_o,
                   "flops_per_row_fwd": fo_fwd, "log_prob_tflops": o_lp / world * fo_fwd / 1e12}
            if name == "c5":
                engo.zero_grads()
                r_fit, ms_fit = time_pass(lambda: engo.backward(Xo, Co, None, per_o, -1.0 / per_o), per_o, reps=3)
                ent.update({"fit_kernel_rows_s": r_fit, "fit_rows_per_launch": per_o, "flops_per_row_fit": fo_fit,
                            "fit_tflops": r_fit / world * fo_fit / 1e12,
                            "fit_kernel": "rnvp_tile_kernel<TR,2> (FP32-FMA; no tensor-core path for D=128 / H=512 yet)"})
            others[name + " -- " + desco] = ent
            del Xo, Co, engo, nfo
        if world == 1:      # single process only: RealNVP.fit under a process group is collective (all ranks would have to join)
            try:
                from sklearn.datasets import make_moons
                Xm, ym = make_moons(n_samples=1000, noise=0.1, random_state=0)
                torch.manual_seed(0)
                mm = RealNVP(lr=0.01, n_epochs=100)
                mm.fit(Xm[:64], ym[:64].reshape(-1, 1))                 # lazy init + first launches
                mm.loss_history.clear()
                t0 = time.perf_counter()
                mm.fit(Xm, ym.reshape(-1, 1))
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                t0 = time.perf_counter()
                Sm = mm.sample(ym.reshape(-1, 1))
                ds = time.perf_counter() - t0
                others["c1 -- configs[0]: README make_moons RealNVP(lr=0.01, n_epochs=100), 1000 rows, batch 32"] = {
                    "fit_wall_s": dt, "steps": len(mm.loss_history), "rows_per_s": 100000 / dt,
                    "us_per_step": dt / max(len(mm.loss_history), 1) * 1e6, "final_loss": float(mm.loss_history[-1]),
                    "sample_1000_rows_ms": ds * 1e3, "sample_shape": list(Sm.shape),
                    "note": "through the public API; launch-bound (3 launches per 32-row step); reference CPU: 43 s (SURVEY 6)"}
            except Exception as e:                                       # sklearn missing etc.: report, do not fail the bench
                others["c1"] = {"skipped": repr(e)}

    # ---- end to end through the public API with pinned host arrays
    e2e = None
    if not args.no_e2e:
        model = RealNVP(n_layers=L, hidden=hidden, activation="tanh", batch_size=n_global, n_epochs=1, lr=lr)
        n_e2e = n_global * max(4, min(K, 16, 32 // world))          # bounded host memory: every rank holds the whole set
        hgen = torch.Generator().manual_seed(7)                      # same host data on every rank
        Xh = torch.randn(n_e2e, D, generator=hgen).pin_memory()
        Ch = torch.randn(n_e2e, Cd, generator=hgen).pin_memory() if Cd else None
        torch.manual_seed(0)
        model.fit(Xh, Ch)      # warm-up with the same shapes: lazy init, workspace and pinned staging buffers, first launches
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.fit(Xh, Ch)                                           # H2D of all rows, steps, D2H of the losses
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        steps_e2e = n_e2e // n_global
        # the same call with the opt-in GPU shuffle (not the reference's batch composition): shows what the sequential
        # CPU shuffle of the reference-faithful default costs once several GPUs share one global batch
        model.shuffle = "device"
        model.fit(Xh, Ch)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.fit(Xh, Ch)
        torch.cuda.synchronize()
        dt2 = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt2, op=dist.ReduceOp.MAX)
        # sample() end to end: conditions from pinned host memory in, numpy rows out (H2D of C, randn + inverse kernel, D2H)
        n_smp = min(n_e2e, 1 << 20)
        model.sample(Ch[:4096] if Ch is not None else 4096)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Xs = model.sample(Ch[:n_smp] if Ch is not None else n_smp)
        dts = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dts, op=dist.ReduceOp.MAX)
        e2e_sample = {"value": n_smp * world / float(dts), "unit": "rows/s", "rows_per_call": n_smp,
                      "h2d_bytes_per_call": n_smp * 4 * Cd, "d2h_bytes_per_call": int(Xs.nbytes),
                      "api": "RealNVP.sample(C_host) -> numpy (every rank samples its own rows, no communication)"}
        del Xs
        e2e = {"value": n_e2e / float(dt), "unit": "rows/s",
               "h2d_bytes_per_step": n_global * 4 * (D + Cd) + 8 * n_global, "d2h_bytes_per_step": 4,
               "steps": steps_e2e, "api": "RealNVP.fit(X_host, C_host), n_epochs=1, replicated data-parallel",
               "shuffle": "reference (default): batches composed exactly as the reference's DataLoader does",
               "value_with_device_shuffle": n_e2e / float(dt2), "sample": e2e_sample}

    if rank == 0:
        H = hidden[0]
        f_fwd, f_fit = flops_per_row(D, Cd, L, H)
        kms = float(kern_ms)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        fp32_peak = 2 * 128 * sms * sm_max * 1e6 / 1e12                   # TFLOP/s, FFMA pipe
        achieved = per_gpu * f_fit / (kms * 1e-3) / 1e12
        bytes_row = 4 * (D + Cd) + 8
        line = {
            "metric": "RealNVP fit rows/sec (fwd+bwd+Adam)", "value": rows_per_s, "unit": "rows/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload + " -- " + desc, "D": D, "Cd": Cd, "n_layers": L,
                       "hidden": list(hidden), "activation": "tanh", "rows_per_gpu_per_step": per_gpu,
                       "global_batch": n_global, "parallelism": f"dp{world}" if world > 1 else "single",
                       "l2": f"inputs larger than L2: each step gathers a fresh random batch from a resident "
                             f"{n_res * 4 * (D + Cd) >> 20} MiB data set",
                       "final_loss": final_loss},
            "roofline": {"bound": "fp32_fma", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp32_peak, "traffic": None,
                         "kernel": ("rnvp_mma_kernel<..,2> (tcgen05 TF32x3 forward + backward sweeps) + rnvp_wgrad_kernel "
                                    "(mma.sync TF32x3 weight-gradient sweep), timed together"
                                    if eng.fit_on_tensor_cores else
                                    "rnvp_mma_kernel<..,2> (tcgen05 forward sweep) + rnvp_tile_kernel<TR,3> (FP32 backward sweep)"
                                    if eng._bwd_two_kernels else "rnvp_tile_kernel<TR,2> (fused forward+backward)"),
                         "kernel_ms": kms,
                         "kernel_share_of_step": kms / (total_ms / K),
                         "flops_per_row": f_fit, "rows_per_launch": per_gpu,
                         "peak_source": f"FP32-FMA pipe: 2*128 lanes*{sms} SMs*{sm_max:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz) -- "
                                        "the roofline of the reference's arithmetic type; the TF32x3 tensor-core kernels spend 3 "
                                        "MMA flops per algorithmic flop and are bound by tcgen05.mma issue + the tanh epilogue (DESIGN.md 4)",
                         "hbm_gbs": per_gpu * bytes_row / (kms * 1e-3) / 1e9,
                         "hbm_frac_of_measured": per_gpu * bytes_row / (kms * 1e-3) / 1e9 / hbm_peak},
            "gpu_launches": launches, "clocks": clocks,
        }
        bf16_sust = float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1346.6)
        if eng.fit_on_tensor_cores:
            # transparency: what the tensor pipes actually execute (3 TF32 MMA passes per algorithmic MAC, operand padding
            # not counted) against the dense TF32 rate (half the measured bf16 rate) -- they are far from saturated; and
            # the DRAM bytes of one step from the committed ncu capture of this command's kernels (activation records)
            line["roofline"]["tensor_pipe"] = {
                "executed_tf32_tflops": 3 * achieved, "peak_tf32_tflops": bf16_sust / 2,
                "frac": 3 * achieved / (bf16_sust / 2),
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 runs at half the bf16 rate)"}
            if per_gpu == 75776 and args.workload == "c3":
                line["roofline"]["traffic"] = 4.52e9
                line["roofline"]["traffic_note"] = ("dram__bytes_read+write per step from profiles/r01_g_ncu_full_c3_fit_raw.csv: "
                                                    "tcgen05 kernel 1.33 + 1.68 GB, weight-gradient sweep 1.51 GB (activation "
                                                    "records written once, read twice, by design; rows themselves are 12.7 MB)")
        if wgrad_ms is not None:
            f_wgrad = sum(2 * (2 * H * ((D - (i & 1) + 1) // 2) + 2 * H * (D - (D - (i & 1) + 1) // 2 + Cd)) for i in range(L))
            rec_bytes = npad * L * eng.lib.rnvp_wgrad_record_floats(eng._desc) * 4
            line["roofline"]["kernels"] = {
                "rnvp_mma_kernel<16,8,32,0,1,2>": {
                    "ms": kms - wgrad_ms, "algorithmic_flops_per_row": f_fit - f_wgrad,
                    "tflops": per_gpu * (f_fit - f_wgrad) / ((kms - wgrad_ms) * 1e-3) / 1e12,
                    "hbm_bytes_per_launch": per_gpu * bytes_row + per_gpu * L * D * 8 + rec_bytes,
                    "note": "forward sweep + backward sweep (recompute, dgrad); writes the activation records"},
                "rnvp_wgrad_kernel<3,2>": {
                    "ms": wgrad_ms, "algorithmic_flops_per_row": f_wgrad,
                    "tflops": per_gpu * f_wgrad / (wgrad_ms * 1e-3) / 1e12,
                    "hbm_bytes_per_launch": rec_bytes, "hbm_gbs": rec_bytes / (wgrad_ms * 1e-3) / 1e9,
                    "hbm_frac_of_measured": rec_bytes / (wgrad_ms * 1e-3) / 1e9 / hbm_peak,
                    "note": "timed alone on the last step's records"}}
        mufu_peak = 16 * sms * sm_max * 1e6                             # MUFU lanes/s: the tanh (ex2 + rcp) pipe
        n_tanh = 2 * H * L
        line["phases"] = {
            "log_prob": {"value": lp_rate, "unit": "rows/s", "kernel": fam, "rows_per_launch": n_pass, "kernel_ms": lp_ms,
                         "flops_per_row": f_fwd, "frac_of_fp32_fma_peak": lp_rate / world * f_fwd / 1e12 / fp32_peak,
                         "frac_of_mufu_peak": lp_rate / world * 2 * n_tanh / mufu_peak},
            "sample": {"value": s_rate, "unit": "rows/s", "kernel": fam, "rows_per_launch": n_pass, "kernel_ms": s_ms,
                       "flops_per_row": f_fwd, "frac_of_fp32_fma_peak": s_rate / world * f_fwd / 1e12 / fp32_peak,
                       "frac_of_mufu_peak": s_rate / world * 2 * n_tanh / mufu_peak,
                       "note": "latent noise read from HBM (parity mode)"},
        }
        f2_fwd, _ = flops_per_row(D2, Cd2, L2, hid2[0])
        line["also"] = {"c2 -- " + desc2: {
            "log_prob_rows_s": c2_lp, "sample_rows_s": c2_s, "rows_per_launch": n2, "kernel": "small-flow kernel (row per thread)",
            "frac_of_mufu_peak": c2_lp / world * 2 * (2 * hid2[0] * L2) / mufu_peak,
            "frac_of_fp32_fma_peak": c2_lp / world * f2_fwd / 1e12 / fp32_peak,
            "hbm_gbs": c2_lp / world * 16 / 1e9}}
        line["also"].update(others)
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            rows = 16384
            rate, cores, step_s = cpu_port_step_rate(D, Cd, L, hidden, rows, reps=3, warm=1)
            line["cpu_baseline"] = {"value": rate, "unit": "rows/s", "cores": cores, "kind": "port",
                                    "sample": f"3 optimisation steps of {rows} rows ({step_s:.2f} s each), oracle port"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
