#!/bin/bash
# round 2: persistent double-buffered pair kernel: parity, timing vs the one-block-per-chunk kernel, ncu
mkdir -p gpurun_out
PARM_B200_TILE_CHECK=1 timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -4
python - <<'PY' 2>gpurun_out/r2e.err | tee gpurun_out/r2e_sweep.jsonl
import os, sys, json
sys.path.insert(0, os.getcwd())
sys.argv = ["tile_sweep"]
import tools.tile_sweep as ts
from parm_b200 import workloads as W
w = W.config3(100)
KEYS = ("PARM_B200_TILE_BANKS", "PARM_B200_BUILD_MASKS", "PARM_B200_TILE_STAGE", "PARM_B200_K1_PREL", "PARM_B200_TILE_PERS")
for env in [{"PARM_B200_TILE_PERS": 0}, {"PARM_B200_TILE_PERS": 1}, {"PARM_B200_TILE_PERS": 1, "PARM_B200_TILE_BANKS": 1},
            {"PARM_B200_TILE_PERS": 1, "PARM_B200_BUILD_MASKS": 1}]:
    for k in KEYS:
        os.environ.pop(k, None)
    e = {"PARM_B200_TILE": 1}
    e.update(env)
    ts.run(w, 200, e)
PY
tail -3 gpurun_out/r2e.err
for b in 0; do
PARM_B200_TILE_BANKS=$b ncu --set full --clock-control none --import-source on -k regex:'k_force_tile|k_tile_localize' -s 4 -c 2 \
    -o gpurun_out/r2e_tile_b$b -f python tools/tile_probe.py --steps 10 > gpurun_out/r2e_ncu_b$b.log 2>&1
tail -2 gpurun_out/r2e_ncu_b$b.log
ncu -i gpurun_out/r2e_tile_b$b.ncu-rep --page source --csv > gpurun_out/r2e_tile_b${b}_source.csv 2>/dev/null
ncu -i gpurun_out/r2e_tile_b$b.ncu-rep --page raw --csv > gpurun_out/r2e_tile_b${b}_raw.csv 2>/dev/null
rm -f gpurun_out/r2e_tile_b$b.ncu-rep
done
