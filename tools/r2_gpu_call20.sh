#!/bin/bash
# k_huffman<64> with two units per thread (big_values loops side by side): parity, then timing at 24 warps (80 regs, spills) / 20 warps (96 regs)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests20.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests20.log
run() { name=$1; shift
  for wl in cfg3 cfg4; do
    env "$@" timeout 300 python tools/profile_run.py --streams 2048 --passes 3 --workload $wl > gpurun_out/r2n_${wl}_$name.log 2>&1
    echo "$wl $name $(tail -n 1 gpurun_out/r2n_${wl}_$name.log | cut -c1-120)"
  done
}
run d768 X=1
run d640 MP3GPU_LIB_VARIANT=d640
run d768_upw64 MP3GPU_K1_UPW=64
run d640_upw64 MP3GPU_LIB_VARIANT=d640 MP3GPU_K1_UPW=64
