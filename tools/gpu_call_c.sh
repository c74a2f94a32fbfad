#!/bin/bash
tag=${1:-r02_d}
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_fit_parity.py tests/test_gpu_mma.py -m gpu -q -x > $out/${tag}_tests.log 2>&1; tail -3 $out/${tag}_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err; python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench.json"))
    k = d["roofline"].get("kernels") or {}
    print("bench", d["value"], d["ms_per_step"], {n: v["ms"] for n, v in k.items()}, "e2e", d["e2e"]["value"], d["e2e"]["value_with_device_shuffle"], d["e2e"]["sample"]["value"])
    print("phases", d["phases"]["log_prob"]["value"], d["phases"]["sample"]["value"])
    for n, v in d["also"].items(): print(n[:3], {a: b for a, b in v.items() if isinstance(b, (int, float))})
except Exception as e:
    print("bench failed:", e)
PY
tail -3 $out/${tag}_bench.err
