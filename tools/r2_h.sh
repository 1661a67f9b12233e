#!/bin/bash
# full GPU suite with the round-2 defaults, then the 1-GPU bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2h_pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench_1gpu.json 2> gpurun_out/r2h_bench_1gpu.err; tail -2 gpurun_out/r2h_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2h_bench_1gpu.json').read().strip().splitlines()[-1])
print('value %.4g ms/step %.4f e2e %.4g frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
print(d['roofline']['step_share']); print(d['steady_state'])
PY
