#!/bin/bash
mkdir -p gpurun_out
python tools/small_n_probe.py --sides 2>>gpurun_out/r2w.err | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 300 -c 60 --csv --log-file gpurun_out/r2w_launches_config1.csv python tools/small_n_probe.py --steps 200 --sides > gpurun_out/r2w.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2w_launches_config1.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
for k, v in agg.items(): print('%-72s n=%3d mean %.2f us' % (k, len(v), sum(v) / len(v) / 1e3 if max(v) > 500 else sum(v)/len(v)))
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_tests.py tests/test_gpu_integrators.py -m gpu -x -q 2>&1 | tail -2
python tools/pair_variants.py --steps 240 "" 2>>gpurun_out/r2w.err | cut -c1-500
