#!/bin/bash
mkdir -p gpurun_out
# parity of the sub-wave (L2 hand-off) mode: the stage-level and batch-level suites with tiny and realistic sub-waves
MP3GPU_SUB=1000 python -m pytest tests/test_gpu_scale.py tests/test_gpu_synth.py tests/test_gpu_fixtures.py tests/test_gpu_ranges.py tests/test_gpu_multidevice.py -m gpu -x -q > gpurun_out/r2_tests7_sub1000.log 2>&1; echo "sub1000 rc=$?"; tail -2 gpurun_out/r2_tests7_sub1000.log
MP3GPU_SUB=16384 python -m pytest tests/test_gpu_scale.py tests/test_gpu_decoder_api.py -m gpu -x -q > gpurun_out/r2_tests7_sub16384.log 2>&1; echo "sub16384 rc=$?"; tail -2 gpurun_out/r2_tests7_sub16384.log
run() { name=$1; shift
  for wl in cfg3 cfg4; do
    env "$@" timeout 300 python tools/profile_run.py --streams 2048 --passes 3 --workload $wl > gpurun_out/r2e_${wl}_$name.log 2>&1
    echo "$wl $name $(tail -n 1 gpurun_out/r2e_${wl}_$name.log | cut -c1-140)"
  done
}
run base X=1
run sub8k_s8_y6 MP3GPU_SUB=8192 MP3GPU_SUB_SEG=8 MP3GPU_SUB_SYN=6
run sub12k_s8_y6 MP3GPU_SUB=12288 MP3GPU_SUB_SEG=8 MP3GPU_SUB_SYN=6
run sub16k_s8_y6 MP3GPU_SUB=16384 MP3GPU_SUB_SEG=8 MP3GPU_SUB_SYN=6
run sub16k_s4_y6 MP3GPU_SUB=16384 MP3GPU_SUB_SEG=4 MP3GPU_SUB_SYN=6
run sub16k_s8_y4 MP3GPU_SUB=16384 MP3GPU_SUB_SEG=8 MP3GPU_SUB_SYN=4
run sub16k_s8_y12 MP3GPU_SUB=16384 MP3GPU_SUB_SEG=8 MP3GPU_SUB_SYN=12
run sub16k_s16_y8 MP3GPU_SUB=16384 MP3GPU_SUB_SEG=16 MP3GPU_SUB_SYN=8
run sub20k_s8_y6 MP3GPU_SUB=20480 MP3GPU_SUB_SEG=8 MP3GPU_SUB_SYN=6
run sub24k_s8_y6 MP3GPU_SUB=24576 MP3GPU_SUB_SEG=8 MP3GPU_SUB_SYN=6
python tools/e2e_probe.py > gpurun_out/r2_e2e_probe_default.log 2>&1; tail -n 1 gpurun_out/r2_e2e_probe_default.log
timeout 300 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_huffman -c 1 -o gpurun_out/r2_k1_v5_cfg3 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r2_ncu_k1a.log 2>&1
timeout 300 python tools/profile_run.py --streams 2048 --passes 1 --workload cfg4 > gpurun_out/r2_ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_huffman -c 1 -o gpurun_out/r2_k1_v5_cfg4 python tools/profile_run.py --streams 2048 --passes 1 --workload cfg4 > gpurun_out/r2_ncu_k1b.log 2>&1
