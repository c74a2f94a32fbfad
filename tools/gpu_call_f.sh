#!/bin/bash
tag=${1:-r02_k}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_fit_parity.py -m gpu -q -x -k "c5 or wide or c4" > $out/${tag}_widefit.log 2>&1; tail -25 $out/${tag}_widefit.log
