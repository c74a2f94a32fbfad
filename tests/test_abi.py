"""The C-ABI library loads without a GPU and exports every symbol include/rnvp.h declares."""
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rnvp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rnvp_[a-z_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from probaforms_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 14
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    lib = _lib.load()                       # dlopen only: no CUDA call is made
    for n in names:
        assert hasattr(lib, n), n
    assert lib.rnvp_version() >= 100
    assert lib.rnvp_last_error() is not None


def test_no_cpu_fallback_in_product():
    """The product never imports the oracle and refuses to run without CUDA."""
    import pytest
    import torch
    pkg = os.path.join(ROOT, "probaforms_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU", ""), f
    if not torch.cuda.is_available():
        from probaforms_b200.models import RealNVP
        import numpy as np
        with pytest.raises(RuntimeError):
            RealNVP().fit(np.zeros((8, 2)), None)
