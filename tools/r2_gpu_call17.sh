#!/bin/bash
# k_synth1 (one warp per segment and channel): parity, then timing against k_synth at 4 / 5 / 6 CTAs per SM
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests17.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests17.log
run() { name=$1; shift
  for wl in cfg3 cfg4; do
    env "$@" timeout 300 python tools/profile_run.py --streams 2048 --passes 3 --workload $wl > gpurun_out/r2s_${wl}_$name.log 2>&1
    echo "$wl $name $(tail -n 1 gpurun_out/r2s_${wl}_$name.log | cut -c1-200)"
  done
}
run old MP3GPU_SYNTH1=0
run s5 X=1
run s4 MP3GPU_LIB_VARIANT=s4
run s6 MP3GPU_LIB_VARIANT=s6
