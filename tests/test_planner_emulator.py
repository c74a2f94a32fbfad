"""CPU check of the host planner: run the per-tile op programs in the test-only host emulator
(tests/emul/rnvp_emul.cpp, poisoned shared memory) and compare with the reference's goldens.

This validates packed layout maps, smem carve-up, chunking, stash and flags of the programs the
CUDA kernel will execute; the device micro-kernels themselves are covered by the -m gpu tests.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, SMALL_CASES, SEEDED_CASES, load_golden, golden_params, rel_err
from oracle import realnvp_oracle as O

EMUL_DIR = os.path.join(ROOT, "tests", "emul")
TOL = 1e-5      # fp32 tolerance of north_star (max-abs error / max-abs value)


@pytest.fixture(scope="session")
def emul():
    so = os.path.join(EMUL_DIR, "librnvp_emul.so")
    src = os.path.join(EMUL_DIR, "rnvp_emul.cpp")
    hdr = os.path.join(ROOT, "probaforms_b200", "csrc", "rnvp_planner.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    lib.rnvp_emulate.restype = C.c_int
    return lib


def fptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run(lib, mode, D, Cd, L, hidden, act, flat, X, Cn, N, l0=0, l1=None, scale=1.0, TR=0, idx=None):
    l1 = L if l1 is None else l1
    hid = (C.c_int * len(hidden))(*hidden)
    out_x = np.full((N, D), np.nan, np.float32)
    ld = np.full(N, np.nan, np.float32)
    lp = np.full(N, np.nan, np.float32)
    gflat = np.zeros(flat.size, np.float32)
    loss = C.c_double(0)
    info = (C.c_int * 5)()
    rc = lib.rnvp_emulate(mode, D, Cd, L, len(hidden), hid, 1 if act == "tanh" else 2, TR,
                          fptr(flat), fptr(X), fptr(Cn), fptr(idx), C.c_longlong(N), l0, l1, C.c_float(scale),
                          fptr(out_x), fptr(ld), fptr(lp), fptr(gflat), C.byref(loss), info)
    assert rc == 0, rc
    return out_x, ld, lp, gflat, loss.value, list(info)


def flat_of(params, L, nh):
    return np.concatenate([params[k].numpy().reshape(-1) for k in O.param_order(L, nh)]).astype(np.float32)


def cfg(g):
    return int(g["D"]), int(g["Cd"]), int(g["L"]), tuple(int(h) for h in g["hidden"]), str(g["activation"])


@pytest.mark.parametrize("TR", [8, 4, 2])
@pytest.mark.parametrize("name", SMALL_CASES)
def test_programs_small_cases(emul, name, TR):
    g = load_golden(name)
    D, Cd, L, hidden, act = cfg(g)
    p = golden_params(g)
    flat = flat_of(p, L, len(hidden))
    X = np.ascontiguousarray(g["X"])
    Cn = np.ascontiguousarray(g["C"]) if "C" in g else None
    N = X.shape[0]
    z, ld, lp, _, _, info = run(emul, 0, D, Cd, L, hidden, act, flat, X, Cn, N, TR=TR)
    assert rel_err(z, g["z"]) < TOL and rel_err(ld, g["logdet"]) < TOL and rel_err(lp, g["logp"]) < TOL
    x, *_ = run(emul, 1, D, Cd, L, hidden, act, flat, np.ascontiguousarray(g["eps"]), Cn, N, TR=TR)
    assert rel_err(x, g["sample"]) < TOL
    # single layer f / g (layer 1 has the odd mask)
    y1, ld1, *_ = run(emul, 0, D, Cd, L, hidden, act, flat, X, Cn, N, l0=1, l1=2, TR=TR)
    if L > 1:
        assert rel_err(y1, g["layer1_f"]) < TOL
        assert np.max(np.abs(ld1 - g["layer1_logdet"])) < TOL * max(1.0, np.max(np.abs(g["layer1_logdet"])))
        x1, *_ = run(emul, 1, D, Cd, L, hidden, act, flat, X, Cn, N, l0=1, l1=2, TR=TR)
        assert rel_err(x1, g["layer1_g"]) < TOL
    # fused forward+backward: loss = -mean logp  -> scale = -1/N
    _, _, lp2, gflat, loss_sum, _ = run(emul, 2, D, Cd, L, hidden, act, flat, X, Cn, N, scale=-1.0 / N, TR=TR)
    assert rel_err(lp2, g["logp"]) < TOL
    assert abs(-loss_sum / N - float(g["loss"])) < TOL * abs(float(g["loss"]))
    gref = np.concatenate([g["g/" + k].reshape(-1) for k in O.param_order(L, len(hidden))])
    assert rel_err(gflat, gref) < 2e-5
    assert np.all(gflat[gref == 0] == 0) or name in ("multi_hidden_relu", "unknown_act")   # masked entries exact 0


@pytest.mark.parametrize("name", SEEDED_CASES)
def test_programs_bench_shapes(emul, name):
    g = load_golden(name)
    D, Cd, L, hidden, act = cfg(g)
    seed, N = int(g["seed"]), int(g["N"])
    p = O.init_params(D, Cd, L, hidden, seed=seed)
    flat = flat_of(p, L, len(hidden))
    gen = torch.Generator().manual_seed(seed + 1000)
    X = torch.randn(N, D, generator=gen).numpy()
    Cn = torch.randn(N, Cd, generator=gen).numpy()
    eps = torch.randn(N, D, generator=gen).numpy()
    z, ld, lp, _, _, info = run(emul, 0, D, Cd, L, hidden, act, flat, X, Cn, N)
    assert rel_err(z, g["z"]) < TOL and rel_err(lp, g["logp"]) < TOL
    x, *_ = run(emul, 1, D, Cd, L, hidden, act, flat, eps, Cn, N)
    assert rel_err(x, g["sample"]) < TOL
    _, _, _, gflat, loss_sum, info2 = run(emul, 2, D, Cd, L, hidden, act, flat, X, Cn, N, scale=-1.0 / N)
    assert abs(-loss_sum / N - float(g["loss"])) < TOL * abs(float(g["loss"]))
    err = np.max(np.abs(gflat[g["grad_idx"]] - g["grad_vals"])) / float(g["grad_absmax"])
    assert err < 2e-5
    assert int((gflat != 0).sum()) <= int(g["grad_nnz"]) + 8
    print(name, "fwd plan", info, "bwd plan", info2)


def test_row_gather_and_ragged_tail(emul):
    g = load_golden("t5c3_tanh")
    D, Cd, L, hidden, act = cfg(g)
    flat = flat_of(golden_params(g), L, len(hidden))
    X, Cn = np.ascontiguousarray(g["X"]), np.ascontiguousarray(g["C"])
    idx = np.random.RandomState(0).permutation(X.shape[0])[:37].astype(np.int64)
    z, ld, lp, *_ = run(emul, 0, D, Cd, L, hidden, act, flat, X, Cn, 37, idx=idx)
    assert rel_err(lp, g["logp"][idx]) < TOL and rel_err(z, g["z"][idx]) < TOL
