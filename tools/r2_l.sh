#!/bin/bash
# ncu launch list of the bench command (per-launch durations, cold cache, serialised)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 30 --warmup 3 --equil 100 --steady-steps 30 --no-cpu > gpurun_out/r2l_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r2l_bench_under_ncu.log | cut -c1-200
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2l_launches.csv', errors='ignore')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    name = r[kn].split('(')[0][:60]
    tot[name][0] += 1; tot[name][1] += v
unit = rows[hi+1][hdr.index('Metric Unit')] if len(rows) > hi+1 else '?'
s = sum(v[1] for v in tot.values())
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:24]:
    print('%-62s n=%4d total %10.1f %s  avg %9.2f  share %5.1f%%' % (k, c, t, unit, t / c, 100 * t / s))
PY
