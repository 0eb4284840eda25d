#!/bin/bash
# k_hybrid: one register path for long, short and mixed blocks (short-block reorder as a per-lane gather)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests19.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests19.log
for wl in cfg3 cfg4; do
  timeout 300 python tools/profile_run.py --streams 2048 --passes 3 --workload $wl > gpurun_out/r2m_${wl}.log 2>&1
  echo "$wl $(tail -n 1 gpurun_out/r2m_${wl}.log | cut -c1-300)"
done
