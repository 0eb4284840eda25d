#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests3.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests3.log
run() { # name env... 
  name=$1; shift
  for wl in cfg3 cfg4; do
    env "$@" timeout 300 python tools/profile_run.py --streams 4096 --passes 3 --workload $wl > gpurun_out/r2c_k1_${wl}_$name.log 2>&1
  done
}
run default X=1
run upw32 MP3GPU_K1_UPW=32
run upw64 MP3GPU_K1_UPW=64
run upw64_w16 MP3GPU_K1_UPW=64 MP3GPU_K1_WARPS=16
run upw64_p125 MP3GPU_K1_UPW=64 MP3GPU_K1_STAGE_PCT=125
run upw32_p125 MP3GPU_K1_UPW=32 MP3GPU_K1_STAGE_PCT=125
timeout 300 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_huffman -c 1 -o gpurun_out/r2_k1_v3 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_k1.log 2>&1
tail -3 gpurun_out/r2_tests3.log; tail -q -n 1 gpurun_out/r2c_k1_*.log | cut -c1-60; ls gpurun_out/r2c_k1_*.log
