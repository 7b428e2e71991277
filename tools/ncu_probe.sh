#!/bin/bash
# usage: tools/ncu_probe.sh B tag   -- one ncu --set full capture of the solver kernel (second launch) at batch size B
B=${1:-32768}; TAG=${2:-probe}
python tools/gpu_probe.py $B f64 1 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    python tools/gpu_probe.py $B f64 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/plain_$TAG.log
