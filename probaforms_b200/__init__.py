"""B200-native RealNVP hot path, drop-in for ``probaforms.models.RealNVP``.

Only the RealNVP path of hse-cs/probaforms is provided:
``from probaforms_b200.models import RealNVP``.  All arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI in ``include/rnvp.h``
(``csrc/librnvp_b200.so``); there is no CPU path and no fallback.
"""
__version__ = "0.1.0"
