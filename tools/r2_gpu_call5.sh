#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests5.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests5.log
tail -3 gpurun_out/r2_tests5.log
( time python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err ) 2> gpurun_out/r2_bench_n1.time
echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.time
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>&1
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('value',d['value'],'ms',d['ms_per_step'],{k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
print('e2e',json.dumps(d['e2e'])[:900])
print('cpu',d['cpu_baseline'])
print('parity',d['parity'])
c=d['cfg4']; print('cfg4',c['value'],c['ms_per_step'],{k:round(v['ms_per_step'],2) for k,v in c['kernels'].items()}, c['parity'], c.get('cpu_baseline'), c.get('x_cpu_baseline_device_resident'), c.get('x_cpu_baseline_e2e'))
print('cfg4 e2e', json.dumps(c['e2e'])[:600])
print('cfg5',json.dumps(d['cfg5'])[:1500])
"
