#!/bin/bash
# full GPU test suite, sanitizers on the new kernels, full bench
tag=${1:-r02_l}
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -x > $out/${tag}_tests.log 2>&1; tail -4 $out/${tag}_tests.log
for tool in racecheck synccheck memcheck; do
  for shape in "32 8 2 64 300" "128 32 2 128 300" "64 16 2 128 200"; do
    timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/dbg_bwd.py $shape > $out/${tag}_${tool}_$(echo $shape | tr ' ' '_').log 2>&1
    echo "$tool [$shape] rc=$? $(grep -E 'SUMMARY' $out/${tag}_${tool}_$(echo $shape | tr ' ' '_').log | tail -1) $(grep -E 'worst rel' $out/${tag}_${tool}_$(echo $shape | tr ' ' '_').log | tail -1)"
  done
done
python tools/dbg_bwd.py 128 32 2 128 300 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench.json"))
    k = d["roofline"].get("kernels") or {}
    print("bench", d["value"], d["ms_per_step"], d["roofline"]["frac"], {n[:14]: round(v["ms"], 4) for n, v in k.items()})
    print("e2e", d["e2e"]["value"], d["e2e"]["value_with_device_shuffle"], d["e2e"]["sample"]["value"])
    for n, v in d["also"].items():
        if n.startswith("c5") or n.startswith("c4"): print(n[:3], json.dumps(v)[:900])
except Exception as e:
    print("bench failed:", e)
PY
tail -3 $out/${tag}_bench.err
