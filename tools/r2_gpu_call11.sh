#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests11.log 2>&1; echo "tests (group tiles, GW=4) rc=$?"; tail -2 gpurun_out/r2_tests11.log
MP3GPU_K1_GROUP=2 python -m pytest tests/test_gpu_fixtures.py tests/test_gpu_synth.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/r2_tests11_g2.log 2>&1; echo "GW=2 rc=$?"; tail -1 gpurun_out/r2_tests11_g2.log
MP3GPU_K1_GROUP=8 python -m pytest tests/test_gpu_fixtures.py tests/test_gpu_synth.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/r2_tests11_g8.log 2>&1; echo "GW=8 rc=$?"; tail -1 gpurun_out/r2_tests11_g8.log
run() { name=$1; shift
  for wl in cfg3 cfg4; do
    env "$@" timeout 300 python tools/profile_run.py --streams 2048 --passes 3 --workload $wl > gpurun_out/r2g_${wl}_$name.log 2>&1
    echo "$wl $name $(tail -n 1 gpurun_out/r2g_${wl}_$name.log | cut -c1-60)"
  done
}
run g0 MP3GPU_K1_GROUP=0
run g2 MP3GPU_K1_GROUP=2
run g4 MP3GPU_K1_GROUP=4
run g8 MP3GPU_K1_GROUP=8
run g4_p110 MP3GPU_K1_GROUP=4 MP3GPU_K1_STAGE_PCT=110
run g4_p150 MP3GPU_K1_GROUP=4 MP3GPU_K1_STAGE_PCT=150
run g8_p110 MP3GPU_K1_GROUP=8 MP3GPU_K1_STAGE_PCT=110
run g2_p150 MP3GPU_K1_GROUP=2 MP3GPU_K1_STAGE_PCT=150
timeout 300 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_huffman -c 1 -o gpurun_out/r2_k1_v7_cfg3 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_k1a.log 2>&1
timeout 300 python tools/profile_run.py --streams 2048 --passes 1 --workload cfg4 > gpurun_out/r2_ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_huffman -c 1 -o gpurun_out/r2_k1_v7_cfg4 python tools/profile_run.py --streams 2048 --passes 1 --workload cfg4 > gpurun_out/r2_ncu_k1b.log 2>&1
