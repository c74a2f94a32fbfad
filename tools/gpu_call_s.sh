set -x
for w in c3 c4 c5; do timeout 100 python tools/quick_bench.py --workloads $w --passes bwd --reps 20 2>&1 | grep -o "\"workload[^,]*\|\"rows[^,]*\|fit_kernel_Mrows_s[^,]*" | tr '\n' ' '; echo; done
for pair in "c3 4" "c4 3" "c5 2"; do set -- $pair; RNVP_WG_SLICES=$2 timeout 100 python tools/quick_bench.py --workloads $1 --passes bwd --reps 20 2>&1 | grep -o "fit_kernel_Mrows_s[^,]*"; done
timeout 300 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_fit_parity.py tests/test_gpu_mma.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python tools/e2e_timeline.py 100 2>&1 | tail -22
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_s_c1_launches.csv python tools/c1_steps.py > /dev/null 2>&1; tail -12 gpurun_out/r02_s_c1_launches.csv | cut -c1-220
