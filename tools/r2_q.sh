#!/bin/bash
# round-2 evidence, 1 GPU: full GPU suite, default bench line, driver-style bench line, reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2q_pytest_gpu_summary.txt
python bench.py > gpurun_out/r2q_bench_1M.json 2> gpurun_out/r2q_bench.err; tail -2 gpurun_out/r2q_bench.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2q_bench_1M_steps20.json 2>> gpurun_out/r2q_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2q_bench_reference.json 2>> gpurun_out/r2q_bench.err
python - <<'PY'
import json
for f in ('gpurun_out/r2q_bench_1M.json', 'gpurun_out/r2q_bench_1M_steps20.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.4g ms/step %.4f e2e %.4g frac %.3f traffic %s launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches']))
    print('  ', d['roofline']['step_share']); print('  ', d['steady_state']); print('  ', d['clocks'])
d = json.loads(open('gpurun_out/r2q_bench_reference.json').read().strip().splitlines()[-1])
print('reference', d['value'], d['config'].get('n_atoms_timed'))
PY
