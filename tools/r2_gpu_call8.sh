#!/bin/bash
mkdir -p gpurun_out
MP3GPU_K1_MODE=2 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests8_mode2.log 2>&1; echo "mode2 tests rc=$?"; tail -2 gpurun_out/r2_tests8_mode2.log
MP3GPU_SUB=1000 python -m pytest tests/test_gpu_scale.py tests/test_gpu_synth.py tests/test_gpu_fixtures.py tests/test_gpu_ranges.py tests/test_gpu_multidevice.py -m gpu -x -q > gpurun_out/r2_tests8_sub1000.log 2>&1; echo "sub1000 rc=$?"; tail -2 gpurun_out/r2_tests8_sub1000.log
run() { name=$1; shift
  for wl in cfg3 cfg4; do
    env "$@" timeout 300 python tools/profile_run.py --streams 2048 --passes 3 --workload $wl > gpurun_out/r2f_${wl}_$name.log 2>&1
    echo "$wl $name $(tail -n 1 gpurun_out/r2f_${wl}_$name.log | cut -c1-150)"
  done
}
run mode1 X=1
run mode1_p150 MP3GPU_K1_STAGE_PCT=150
run mode2 MP3GPU_K1_MODE=2
run mode2_p130 MP3GPU_K1_MODE=2 MP3GPU_K1_STAGE_PCT=130
run mode2_p200 MP3GPU_K1_MODE=2 MP3GPU_K1_STAGE_PCT=200
run mode2_w24 MP3GPU_K1_MODE=2 MP3GPU_K1_WARPS=24
run sub16k MP3GPU_SUB=16384 MP3GPU_SUB_SEG=8 MP3GPU_SUB_SYN=6
run sub24k MP3GPU_SUB=24576 MP3GPU_SUB_SEG=8 MP3GPU_SUB_SYN=6
run sub12k MP3GPU_SUB=12288 MP3GPU_SUB_SEG=8 MP3GPU_SUB_SYN=6
MP3GPU_K1_MODE=2 timeout 300 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_plain.log 2>&1 && \
MP3GPU_K1_MODE=2 ncu --set full --clock-control none --import-source on -k regex:k_huffman_sorted -c 1 -o gpurun_out/r2_k1_v6_cfg3 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_k1a.log 2>&1
MP3GPU_K1_MODE=2 timeout 300 python tools/profile_run.py --streams 2048 --passes 1 --workload cfg4 > gpurun_out/r2_ncu_plain2.log 2>&1 && \
MP3GPU_K1_MODE=2 ncu --set full --clock-control none --import-source on -k regex:k_huffman_sorted -c 1 -o gpurun_out/r2_k1_v6_cfg4 python tools/profile_run.py --streams 2048 --passes 1 --workload cfg4 > gpurun_out/r2_ncu_k1b.log 2>&1
