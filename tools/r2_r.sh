#!/bin/bash
# 1 GPU bench lines with the aligned window
mkdir -p gpurun_out
python bench.py > gpurun_out/r2r_bench_1M.json 2> gpurun_out/r2r_bench.err; tail -2 gpurun_out/r2r_bench.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2r_bench_1M_steps20.json 2>> gpurun_out/r2r_bench.err
python - <<'PY'
import json
for f in ('gpurun_out/r2r_bench_1M.json', 'gpurun_out/r2r_bench_1M_steps20.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.4g ms/step %.4f e2e %.4g frac %.3f traffic %s launches %d rebuilds %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'], d['rebuilds_in_timed_region']))
    print('  ', d['equilibration']); print('  ', d['steady_state'])
PY
