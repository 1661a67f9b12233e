#!/bin/bash
# round 2: bulk-copy (TMA) staging of the pair kernel: parity, then timing against the classic staging
mkdir -p gpurun_out
PARM_B200_TILE_CHECK=1 timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_scale.py -m gpu -x -q -k "not 10k_steps" 2>&1 | tail -6
python - <<'PY' 2>gpurun_out/r2c.err | tee gpurun_out/r2c_sweep.jsonl
import os, sys, json
sys.path.insert(0, os.getcwd())
sys.argv = ["tile_sweep"]
import tools.tile_sweep as ts
from parm_b200 import workloads as W
w = W.config3(100)
for env in [{"PARM_B200_TILE_STAGE": 0}, {"PARM_B200_TILE_STAGE": 1, "PARM_B200_K1_PREL": 0}, {"PARM_B200_TILE_STAGE": 1},
            {"PARM_B200_TILE_STAGE": 1, "PARM_B200_TILE_BANKS": 1}]:
    for k in ("PARM_B200_TILE_BANKS", "PARM_B200_BUILD_MASKS", "PARM_B200_TILE_STAGE", "PARM_B200_K1_PREL"):
        os.environ.pop(k, None)
    e = {"PARM_B200_TILE": 1}
    e.update(env)
    ts.run(w, 200, e)
PY
tail -3 gpurun_out/r2c.err
