#!/bin/bash
# round 2: K3+K1 fusion: parity, then timing with and without
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py tests/test_gpu_scale.py tests/test_gpu_edge.py tests/test_gpu_golden.py tests/test_gpu_integrators.py tests/test_gpu_trackers.py -m gpu -x -q 2>&1 | tail -5
python - <<'PY' 2>gpurun_out/r2k.err | tee gpurun_out/r2k_sweep.jsonl
import os, sys, json
sys.path.insert(0, os.getcwd())
sys.argv = ["tile_sweep"]
import tools.tile_sweep as ts
from parm_b200 import workloads as W
w = W.config3(100)
KEYS = ("PARM_B200_FUSE_K3K1",)
for env in [{"PARM_B200_FUSE_K3K1": 0}, {"PARM_B200_FUSE_K3K1": 1}]:
    for k in KEYS:
        os.environ.pop(k, None)
    e = {"PARM_B200_TILE": 1}
    e.update(env)
    ts.run(w, 400, e)
PY
tail -3 gpurun_out/r2k.err
