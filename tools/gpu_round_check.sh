#!/bin/bash
# One gpurun call that produces everything a round needs from a single B200 (about 3 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'tools/gpu_round_check.sh rNN_x'
# -> gpurun_out/<tag>_{tests.log,smoke.log,bench.json,launches.csv,fit.ncu-rep,quick.jsonl}
# NEVER wrap a multi-GPU command in a long gpurun --timeout: a hang is charged N x the limit (round 1 lost 84 GPU-minutes
# to one deadlocked 8-GPU bench).  For N > 1 use: gpurun --gpus N --timeout 170 -- 'timeout 140 python -m torch.distributed.run ...'
tag=${1:-check}
out=gpurun_out
mkdir -p $out
timeout 400 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; tail -2 $out/${tag}_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
timeout 300 python bench.py --steps 30 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench.json"))
    print("bench", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("bench failed:", e)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-others > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"rnvp_mma_kernel|rnvp_wgrad_kernel" -s 2 -c 2 \
    -o $out/${tag}_fit -f python tools/quick_bench.py --workloads c3 --rows 75776 --passes bwd --reps 2 > /dev/null 2>&1
timeout 200 python tools/quick_bench.py > $out/${tag}_quick.jsonl 2>&1
ls -la $out | tail -8
