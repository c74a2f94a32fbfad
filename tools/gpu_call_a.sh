#!/bin/bash
# round-2 call A: all GPU tests, sanitizer racecheck / synccheck on the tcgen05 fit path, short bench
tag=${1:-r02_a}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_roundtrip_and_ragged_sizes > $out/${tag}_tests.log 2>&1; tail -5 $out/${tag}_tests.log
timeout 600 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_fit_parity.py -m gpu -q > $out/${tag}_newtests.log 2>&1; tail -15 $out/${tag}_newtests.log
for tool in racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python tools/dbg_bwd.py 32 8 2 64 300 > $out/${tag}_${tool}.log 2>&1
  echo "$tool rc=$?"; tail -3 $out/${tag}_${tool}.log
done
timeout 300 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2
