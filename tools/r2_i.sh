#!/bin/bash
# 2 GPUs: sharded parity (tests/mgpu_check.py, log kept), then the 2-GPU bench line with its parity_check
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/r2i_mgpu_check_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err
tail -3 gpurun_out/r2i_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2i_bench_2gpu.json').read().strip().splitlines()[-1])
print('2gpu value %.4g ms/step %.4f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print(d['parity_check']); print(d['steady_state']); print(d['roofline']['step_share']); print(d['cpu_baseline'])
PY
