#!/bin/bash
# Bound analysis of k_hybrid / k_synth with the probe build (timing only; outputs are wrong by design)
mkdir -p gpurun_out
for p in 0 1 2 3 4 8 12 15; do
  for wl in cfg3; do
    MP3GPU_LIB_VARIANT=probe MP3GPU_PROBE=$p timeout 300 python tools/profile_run.py --streams 2048 --passes 3 --workload $wl > gpurun_out/r2p_${wl}_$p.log 2>&1
    echo "$wl probe=$p $(tail -n 1 gpurun_out/r2p_${wl}_$p.log | cut -c1-200)"
  done
done
