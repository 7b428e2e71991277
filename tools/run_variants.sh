#!/bin/bash
# usage: tools/run_variants.sh TAG "v1 v2 ..." [B] [reps]  -- gpu_probe of build/libigtmpc_<v>.so, results in gpurun_out/var_TAG.log
TAG=$1; VARS=$2; B=${3:-32768}; R=${4:-2}
for v in $VARS; do
  echo "== $v" >> gpurun_out/var_$TAG.log
  IGT_LIB=build/libigtmpc_$v.so python tools/gpu_probe.py $B f64 $R 2>&1 | cut -c1-200 >> gpurun_out/var_$TAG.log
done
cat gpurun_out/var_$TAG.log
