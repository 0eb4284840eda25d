#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests6.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests6.log
tail -3 gpurun_out/r2_tests6.log
( time python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err ) 2> gpurun_out/r2_bench_n1_b.time
echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1_b.err; cat gpurun_out/r2_bench_n1_b.time
for c in 0 4 8 16 32; do MP3HOST_CHUNKS=$c python tools/e2e_probe.py > gpurun_out/r2_e2e_probe_c$c.log 2>&1; tail -n 1 gpurun_out/r2_e2e_probe_c$c.log; done
MP3HOST_CHUNKS=16 python tools/e2e_probe.py --workload cfg4 > gpurun_out/r2_e2e_probe_cfg4_c16.log 2>&1; tail -n 1 gpurun_out/r2_e2e_probe_cfg4_c16.log
# L2 hand-off experiment: small waves so that hyb (4,608 B per granule) stays in the 126 MB L2 between k_hybrid and k_synth
for cfg in "0 32" "65536 32" "65536 16" "32768 16" "32768 8" "16384 8" "16384 4"; do set -- $cfg
  MP3GPU_SEG_LEN=$2 python tools/profile_run.py --streams 1024 --passes 3 --wave $1 > gpurun_out/r2_l2_w$1_s$2.log 2>&1; echo "wave $1 seg $2: $(tail -n 1 gpurun_out/r2_l2_w$1_s$2.log | cut -c1-200)"; done
MP3GPU_SEG_LEN=8 python tools/profile_run.py --streams 256 --passes 1 --wave 16384 > gpurun_out/r2_ncu_plain.log 2>&1 && \
MP3GPU_SEG_LEN=8 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"k_hybrid|k_synth" -s 8 -c 8 --csv --log-file gpurun_out/r2_l2_ncu_w16384_s8.csv python tools/profile_run.py --streams 256 --passes 1 --wave 16384 > gpurun_out/r2_ncu_l2.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n1_b.json'))
print('value',d['value'],'ms',d['ms_per_step'],{k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
print('e2e',d['e2e']['value'], d['e2e'].get('frac_of_copy_ceiling'), d['e2e']['last_call'])
print('output_side', d['output_side'])
print('cfg4 e2e', d['cfg4']['e2e']['value'], d['cfg4']['e2e'].get('frac_of_copy_ceiling'))
print('cfg5 seek', d['cfg5']['seek_to_time'])
"
