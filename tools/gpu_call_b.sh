#!/bin/bash
tag=${1:-r02_b}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_mma.py -m gpu -q -x -k "wgrad or selftest" -s > $out/${tag}_wgrad.log 2>&1; tail -12 $out/${tag}_wgrad.log
timeout 600 python -m pytest tests/test_gpu_fit_parity.py tests/test_gpu_parity.py tests/test_gpu_mma.py -m gpu -q -x --deselect tests/test_gpu_parity.py::test_roundtrip_and_ragged_sizes > $out/${tag}_tests.log 2>&1; tail -8 $out/${tag}_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-others --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench.json"))
    print("bench", d["value"], d["ms_per_step"], d["roofline"].get("kernels"), "e2e", d["e2e"]["value"], d["e2e"]["value_with_device_shuffle"], d["e2e"]["sample"]["value"])
except Exception as e:
    print("bench failed:", e)
PY
tail -3 $out/${tag}_bench.err
RNVP_WGRAD=legacy timeout 200 python bench.py --steps 20 --warmup 5 --no-others --no-cpu-baseline --no-e2e > $out/${tag}_bench_legacy.json 2>/dev/null; python -c "
import json; d=json.load(open('$out/${tag}_bench_legacy.json')); print('legacy', d['value'], d['ms_per_step'], d['roofline'].get('kernels'))"
