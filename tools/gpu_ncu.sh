#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> <kernel regex> <quick_bench args...>: one ncu --set full capture of the named kernel
tag=$1; shift; pat=$1; shift
out=gpurun_out; mkdir -p $out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 2 -c 1 -o $out/${tag} -f python tools/quick_bench.py "$@" > $out/${tag}.log 2>&1
tail -3 $out/${tag}.log; ls -la $out/${tag}.ncu-rep
