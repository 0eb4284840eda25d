#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests2.log
for wl in cfg3 cfg4; do
  for t in 0 256; do
    MP3GPU_K1_THREADS=$t timeout 300 python tools/profile_run.py --streams 4096 --passes 3 --workload $wl > gpurun_out/r2b_k1_${wl}_t$t.log 2>&1
  done
done
timeout 300 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_huffman -c 1 -o gpurun_out/r2_k1_v2 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_k1.log 2>&1
tail -3 gpurun_out/r2_tests2.log; tail -n 1 gpurun_out/r2b_k1_*.log
