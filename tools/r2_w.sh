#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/r2w_launches_config1.csv python tools/small_n_probe.py --steps 200 --sides > gpurun_out/r2w.log 2>&1
tail -2 gpurun_out/r2w.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2w_launches_config1.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
for k, v in agg.items(): print('%-72s n=%3d mean %.2f us' % (k, len(v), sum(v) / len(v) / 1e3 if max(v) > 500 else sum(v)/len(v)))
PY
