#!/bin/bash
# 2 GPUs: one-sort sharded rebuild (default), the two-sort path (PARM_B200_SHARD_FAST=0) and the overflow fall-back
# (message capacity 1 atom) through tests/mgpu_check.py; then the 2-GPU bench line of both paths
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for v in "" "PARM_B200_SHARD_MIGCAP=1" "PARM_B200_SHARD_FAST=0"; do
  echo "== mgpu_check $v" | tee -a gpurun_out/r2u_mgpu_check_2gpu.log
  env $v timeout 600 $TR --master-port 29533 tests/mgpu_check.py 2>&1 | grep -v Warning | grep "mgpu ok\|MGPU\|Error\|error\|assert" | tee -a gpurun_out/r2u_mgpu_check_2gpu.log
done
for v in "" "PARM_B200_SHARD_FAST=0"; do
  env $v timeout 600 $TR --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2u_bench_2gpu_${v: -1}.json 2> gpurun_out/r2u_bench_2gpu.err
  tail -2 gpurun_out/r2u_bench_2gpu.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/r2u_bench_2gpu_${v: -1}.json').read().strip().splitlines()[-1])
print('$v 2gpu value %.4g ms/step %.4f' % (d['value'], d['ms_per_step'])); print(d['parity_check']); print(d['steady_state']); print(d['roofline']['step_share'])
PY
done
