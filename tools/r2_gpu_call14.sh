#!/bin/bash
mkdir -p gpurun_out
for v in r9 r10; do
MP3GPU_LIB_VARIANT=$v python -m pytest tests/test_gpu_fixtures.py tests/test_gpu_synth.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/r2_tests14_$v.log 2>&1; echo "tests $v rc=$?"; tail -1 gpurun_out/r2_tests14_$v.log
done
run() { name=$1; shift
  for wl in cfg3 cfg4; do
    env "$@" timeout 300 python tools/profile_run.py --streams 2048 --passes 3 --workload $wl > gpurun_out/r2j_${wl}_$name.log 2>&1
    echo "$wl $name $(tail -n 1 gpurun_out/r2j_${wl}_$name.log | cut -c1-60)"
  done
}
run r8 X=1
run r9 MP3GPU_LIB_VARIANT=r9
run r10 MP3GPU_LIB_VARIANT=r10
run r10_p110 MP3GPU_LIB_VARIANT=r10 MP3GPU_K1_STAGE_PCT=110
run r9_p110 MP3GPU_LIB_VARIANT=r9 MP3GPU_K1_STAGE_PCT=110
run r10_upw32 MP3GPU_LIB_VARIANT=r10 MP3GPU_K1_UPW=32
