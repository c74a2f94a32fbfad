import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

SMALL_CASES = ["t5c3_tanh", "moons_shape", "d1_regression", "nocond_d5",
               "multi_hidden_relu", "unknown_act", "multi_hidden_tanh"]
SEEDED_CASES = ["c3_shape", "c4_shape", "c5_shape"]
FIT_CASES = ["fit_moons", "fit_nocond_wd"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_params(g, prefix="p/"):
    import torch
    return {k[len(prefix):]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith(prefix)}


def rel_err(a, b):
    """max-abs error normalised by max-abs reference value (SURVEY 8c error metric)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
