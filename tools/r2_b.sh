#!/bin/bash
# round 2: new full-size parity tests, 1-GPU bench line, 2-GPU parity + bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "full_size_1m or 256k" 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err; tail -2 gpurun_out/r2b_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2b_bench_1gpu.json').read().strip().splitlines()[-1])
print('value %.4g ms/step %.4f e2e %.4g frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
print(d['roofline']['step_share']); print(d['steady_state']); print(d['equilibration']); print(d['roofline']['traffic'], d['roofline']['traffic_source'])
PY
if [ "$1" = "2" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2b_bench_2gpu.json 2> gpurun_out/r2b_bench_2gpu.err
tail -3 gpurun_out/r2b_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2b_bench_2gpu.json').read().strip().splitlines()[-1])
print('2gpu value %.4g ms/step %.4f' % (d['value'], d['ms_per_step'])); print(d['parity_check']); print(d['steady_state']); print(d['roofline']['step_share'])
PY
fi
