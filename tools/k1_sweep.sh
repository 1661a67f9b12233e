#!/bin/bash
for g in 1 2 3 4 6; do
  PARM_B200_K1_PER_SM=$g python tools/tile_sweep.py --quick 2>/dev/null | tail -2 | head -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('per_sm $g', 'ms/step %.4f force %.4f rebuild %.4f k1 %.4f k3 %.4f' % (d['ms_per_step'], d['force_ms'], d['rebuild_ms_each'], d['k1_ms'], d['k3_ms']))
"
done
