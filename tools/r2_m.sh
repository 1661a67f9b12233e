#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' 2>gpurun_out/r2m.err | tee gpurun_out/r2m_sweep.jsonl
import os, sys, json
sys.path.insert(0, os.getcwd())
sys.argv = ["tile_sweep"]
import tools.tile_sweep as ts
from parm_b200 import workloads as W
w = W.config3(100)
KEYS = ("PARM_B200_TILE_REGS", "PARM_B200_TILE_CH")
for env in [{}, {"PARM_B200_TILE_REGS": 4, "PARM_B200_TILE_CH": 96}, {"PARM_B200_TILE_REGS": 4, "PARM_B200_TILE_CH": 88}, {"PARM_B200_TILE_REGS": 5, "PARM_B200_TILE_CH": 80}, {"PARM_B200_TILE_REGS": 5, "PARM_B200_TILE_CH": 72}, {"PARM_B200_TILE_CH": 96}]:
    for k in KEYS:
        os.environ.pop(k, None)
    e = {"PARM_B200_TILE": 1}
    e.update(env)
    ts.run(w, 200, e)
PY
tail -3 gpurun_out/r2m.err
