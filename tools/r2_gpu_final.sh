#!/bin/bash
# Final round-2 evidence at N=1: GPU suite, the bench exactly as the driver runs it, the reference arm.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_final.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests_final.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err ) 2> gpurun_out/r2_final_n1.time
echo "bench rc=$?"; cat gpurun_out/r2_final_n1.time | tail -3
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err ) 2> gpurun_out/r2_final_ref.time
echo "ref rc=$?"; tail -3 gpurun_out/r2_final_ref.time
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_final_n1.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],{k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
print('e2e',d['e2e']['value'],d['e2e']['frac_of_copy_ceiling'],d['e2e']['copy_ceiling'])
c=d['cfg4']; print('cfg4',c['value'],c['ms_per_step'],c['e2e']['value'])
c=d['cfg5']; print('cfg5',c['value'],c['ms_per_decode'],c['seek_to_time']['ms_per_seek_and_read'])
r=json.loads([l for l in open('gpurun_out/r2_final_ref.json') if l.startswith('{')][-1]); print('ref',r['value'],r['cpu_baseline']['cores'])
"
