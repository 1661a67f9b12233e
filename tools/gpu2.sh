#!/bin/bash
# 2-GPU check: sharded parity (tests/mgpu_check.py) and the 2-GPU bench line
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -v Warning | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/${1:-bench_2gpu}.json 2> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/${1:-bench_2gpu}.json').read().strip().splitlines()[-1])
print('value %.4g ms/step %.4f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['e2e'], d['roofline']['step_share'])
"
