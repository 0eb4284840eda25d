#!/bin/bash
# N GPUs (run with gpurun --gpus N): parity suite on one GPU, then the bench under torchrun.
mkdir -p gpurun_out
N=${1:-2}
python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests9.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests9.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err ) 2> gpurun_out/r2_bench_n$N.time
echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n$N.err | cut -c1-300; cat gpurun_out/r2_bench_n$N.time
python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2_bench_n$N.json') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],{k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
print('e2e',json.dumps(d['e2e'])[:1000])
c=d['cfg4']; print('cfg4',c['value'],c['ms_per_step'], c.get('x_cpu_baseline_device_resident'), c.get('x_cpu_baseline_e2e'), json.dumps(c['e2e'])[:400])
print('cfg5',json.dumps(d['cfg5'])[:1500])
"
