#!/bin/bash
# full bench + ncu traffic capture + launch list
tag=${1:-r02_j}
out=gpurun_out; mkdir -p $out
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench.json"))
    k = d["roofline"].get("kernels") or {}
    print("bench", d["value"], d["ms_per_step"], d["roofline"]["frac"], {n[:14]: round(v["ms"], 4) for n, v in k.items()})
    print("e2e", d["e2e"]["value"], d["e2e"]["value_with_device_shuffle"], d["e2e"]["sample"]["value"], d["e2e"]["rows"], d["e2e"]["h2d_bytes_per_step"])
    print("phases", {a: b["value"] for a, b in d["phases"].items()})
    for n, v in d["also"].items(): print(n[:3], json.dumps(v)[:700])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench failed:", e)
PY
tail -5 $out/${tag}_bench.err
timeout 300 ncu --set full --clock-control none -k regex:"rnvp_mma_kernel|rnvp_wgrad_tc" -s 2 -c 2 -o $out/${tag}_fit -f python tools/quick_bench.py --workloads c3 --rows 75776 --passes bwd --reps 2 > /dev/null 2>&1
ls -la $out/${tag}_fit.ncu-rep
