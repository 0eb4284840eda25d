#!/bin/bash
# Evidence for profiles/: launch list, full ncu captures of one full wave of every kernel on both workloads, DRAM traffic
# of the sub-wave (L2 hand-off) mode.  Every ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
python tools/profile_run.py --streams 4096 --passes 2 > gpurun_out/r02_plain_4096.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_profile_run_4096.csv python tools/profile_run.py --streams 4096 --passes 2 > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r02_plain_2048.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_huffman|k_hybrid|k_synth" -c 3 -o gpurun_out/r02_full_cfg3 python tools/profile_run.py --streams 2048 --passes 1 > gpurun_out/r02_ncu_full_cfg3.log 2>&1
echo "full cfg3 rc=$?"
python tools/profile_run.py --streams 2048 --passes 1 --workload cfg4 > gpurun_out/r02_plain_2048_cfg4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_huffman|k_hybrid|k_synth" -c 3 -o gpurun_out/r02_full_cfg4 python tools/profile_run.py --streams 2048 --passes 1 --workload cfg4 > gpurun_out/r02_ncu_full_cfg4.log 2>&1
echo "full cfg4 rc=$?"
# DRAM traffic with and without the sub-wave mode, caches left alone and the application replayed (kernel replay would
# flush L2 between the kernels and hide the hand-off)
for sub in 0 16384; do
MP3GPU_SUB=$sub python tools/profile_run.py --streams 256 --passes 1 > gpurun_out/r02_plain_sub$sub.log 2>&1 && \
MP3GPU_SUB=$sub ncu --replay-mode application --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"k_hybrid|k_synth" -c 80 --csv --log-file gpurun_out/r02_dram_sub$sub.csv python tools/profile_run.py --streams 256 --passes 1 > gpurun_out/r02_ncu_sub$sub.log 2>&1
echo "sub $sub rc=$?"
done
python tools/ncu_summary.py gpurun_out/r02_full_cfg3.ncu-rep gpurun_out/r02_ncu_full_summary_cfg3.json gpurun_out/r02_traffic.json
python tools/ncu_summary.py gpurun_out/r02_full_cfg4.ncu-rep gpurun_out/r02_ncu_full_summary_cfg4.json
ls -la gpurun_out/r02_*
