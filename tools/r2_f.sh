#!/bin/bash
# round 2: mask-mode rebuild with the restructured localize pass; persistent kernel tweak
mkdir -p gpurun_out
PARM_B200_BUILD_MASKS=1 PARM_B200_TILE_CHECK=1 timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q -k "not 10k_steps and not full_size_1m" 2>&1 | tail -4
python - <<'PY' 2>gpurun_out/r2f.err | tee gpurun_out/r2f_sweep.jsonl
import os, sys, json
sys.path.insert(0, os.getcwd())
sys.argv = ["tile_sweep"]
import tools.tile_sweep as ts
from parm_b200 import workloads as W
w = W.config3(100)
KEYS = ("PARM_B200_TILE_BANKS", "PARM_B200_BUILD_MASKS", "PARM_B200_TILE_STAGE", "PARM_B200_K1_PREL", "PARM_B200_TILE_PERS", "PARM_B200_BUILD_DIRECT")
for env in [{"PARM_B200_TILE_PERS": 0}, {"PARM_B200_TILE_PERS": 0, "PARM_B200_BUILD_MASKS": 1}]:
    for k in KEYS:
        os.environ.pop(k, None)
    e = {"PARM_B200_TILE": 1}
    e.update(env)
    ts.run(w, 200, e)
PY
tail -3 gpurun_out/r2f.err
PARM_B200_BUILD_MASKS=1 ncu --set full --clock-control none --import-source on -k regex:'k_tile_localize_masks|k_build_cell' -s 2 -c 4 \
    -o gpurun_out/r2f_reb -f python tools/tile_probe.py --steps 12 > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
ncu -i gpurun_out/r2f_reb.ncu-rep --page raw --csv > gpurun_out/r2f_reb_raw.csv 2>/dev/null
rm -f gpurun_out/r2f_reb.ncu-rep
