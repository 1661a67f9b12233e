#!/bin/bash
# 8 GPUs: sharded parity log, then the 8-GPU bench line (weak scaling window + parity_check + config5_16M)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -v Warning | grep -E "mgpu ok|MGPU|Error|error" | tee gpurun_out/r2j_mgpu_check_8gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2j_bench_8gpu.json 2> gpurun_out/r2j_bench_8gpu.err
grep -v Warning gpurun_out/r2j_bench_8gpu.err | tail -5
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2j_bench_8gpu.json').read().strip().splitlines()[-1])
print('8gpu value %.4g ms/step %.4f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print(d['parity_check']); print(d['steady_state']); print(d['roofline']['step_share']); print(d['config5_16M'])
PY
