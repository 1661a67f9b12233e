#!/bin/bash
# 4 GPUs (distinct upper / lower neighbours): sharded parity incl. the migration-path equivalence, then the 4-GPU bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29533 tests/mgpu_check.py 2>&1 | grep -v Warning | grep -E "mgpu ok|MGPU|Error|error|assert|Traceback" | tee gpurun_out/r02_h_mgpu_check_4gpu.log | cut -c1-400
timeout 900 $TR --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_h_bench_4gpu.json 2> gpurun_out/r2y_bench_4gpu.err
grep -v Warning gpurun_out/r2y_bench_4gpu.err | tail -3
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_h_bench_4gpu.json').read().strip().splitlines()[-1])
print('4gpu value %.4g ms/step %.4f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print(d['parity_check']); print(d['steady_state']); print(d['roofline']['step_share'])
PY
