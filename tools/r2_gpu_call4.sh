#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests4.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests4.log
run() { name=$1; shift
  for wl in cfg3 cfg4; do
    env "$@" timeout 300 python tools/profile_run.py --streams 4096 --passes 3 --workload $wl > gpurun_out/r2d_k1_${wl}_$name.log 2>&1
  done
}
run default X=1
run p150 MP3GPU_K1_STAGE_PCT=150
run p300 MP3GPU_K1_STAGE_PCT=300
run w24 MP3GPU_K1_WARPS=24
run w16 MP3GPU_K1_WARPS=16
timeout 300 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_huffman -c 1 -o gpurun_out/r2_k1_v4 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_k1.log 2>&1
tail -3 gpurun_out/r2_tests4.log; for f in gpurun_out/r2d_k1_*.log; do echo "$f $(tail -n 1 $f | cut -c1-50)"; done
