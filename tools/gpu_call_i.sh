#!/bin/bash
out=gpurun_out; mkdir -p $out
for w in 0 1; do
  echo "== RNVP_WIDE16=$w"
  RNVP_WIDE16=$w timeout 300 python tools/quick_bench.py --workloads c3 --rows 1048576 --passes fwd,inv | cut -c300-600
  RNVP_WIDE16=$w timeout 300 python tools/quick_bench.py --workloads c3 --rows 75776 --passes bwd --reps 20 | cut -c300-600
done
RNVP_WIDE16=1 timeout 600 python -m pytest tests/test_gpu_fit_parity.py tests/test_gpu_mma.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
RNVP_WIDE16=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-others --no-cpu-baseline --no-e2e | python -c "
import json,sys; d=json.load(sys.stdin); print('bench wide16', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], {k: v['value'] for k, v in d['phases'].items()})"
