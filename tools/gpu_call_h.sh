#!/bin/bash
tag=${1:-r02_m}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_fit_parity.py tests/test_gpu_mma.py -m gpu -q -x > $out/${tag}_tests.log 2>&1; tail -2 $out/${tag}_tests.log
for tool in synccheck racecheck; do
  for shape in "32 8 2 64 300" "128 32 2 128 300" "64 16 2 128 200"; do
    f=$out/${tag}_${tool}_$(echo $shape | tr ' ' '_').log
    timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/dbg_bwd.py $shape > $f 2>&1
    echo "$tool [$shape] rc=$? $(grep -E 'SUMMARY' $f | tail -1) $(grep -E 'worst rel' $f | tail -1)"
  done
done
