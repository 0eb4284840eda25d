#!/bin/bash
mkdir -p gpurun_out
MP3GPU_K1_STAGED=0 MP3GPU_K1_GROUP=8 python -m pytest tests/test_gpu_fixtures.py tests/test_gpu_synth.py tests/test_gpu_scale.py tests/test_gpu_checked_build.py -m gpu -x -q > gpurun_out/r2_tests13.log 2>&1; echo "tests (unstaged, GW=8) rc=$?"; tail -1 gpurun_out/r2_tests13.log
run() { name=$1; shift
  for wl in cfg3 cfg4; do
    env "$@" timeout 300 python tools/profile_run.py --streams 2048 --passes 3 --workload $wl > gpurun_out/r2i_${wl}_$name.log 2>&1
    echo "$wl $name $(tail -n 1 gpurun_out/r2i_${wl}_$name.log | cut -c1-60)"
  done
}
run c64 MP3GPU_K1_GROUP=0
run u4 MP3GPU_K1_STAGED=0 MP3GPU_K1_GROUP=4
run u8 MP3GPU_K1_STAGED=0 MP3GPU_K1_GROUP=8
run u16 MP3GPU_K1_STAGED=0 MP3GPU_K1_GROUP=16
MP3GPU_K1_STAGED=0 MP3GPU_K1_GROUP=8 timeout 300 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_plain.log 2>&1 && \
MP3GPU_K1_STAGED=0 MP3GPU_K1_GROUP=8 ncu --set full --clock-control none --import-source on -k regex:k_huffman -c 1 -o gpurun_out/r2_k1_v9_cfg3 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_k1a.log 2>&1
