#!/bin/bash
tag=${1:-r02_i}
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_mma.py -m gpu -q -x -k "flow_matches" -s > $out/${tag}_wide.log 2>&1; grep -E "shape \(128|shape \(64, 16, 3|passed|failed|Error|error" $out/${tag}_wide.log | tail -30
timeout 300 python tools/quick_bench.py --workloads c5,c4 --passes fwd,inv > $out/${tag}_quick.jsonl 2>&1; tail -4 $out/${tag}_quick.jsonl | cut -c1-400
