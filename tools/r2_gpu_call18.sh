#!/bin/bash
# K1 prologue latency: cp.async staging behind the sort, tile index one ahead, descriptor prefetch, slen packs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests18.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests18.log
for wl in cfg3 cfg4; do
  timeout 300 python tools/profile_run.py --streams 2048 --passes 3 --workload $wl > gpurun_out/r2l_${wl}.log 2>&1
  echo "$wl $(tail -n 1 gpurun_out/r2l_${wl}.log | cut -c1-300)"
done
