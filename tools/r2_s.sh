#!/bin/bash
# round 2 (second session): pair-kernel prologue (descriptor via shared memory, L2 pull of the next wave's descriptor) and
# warp-uniform passes with a partial last pass (2 / 4 / 6 of 8 entries per lane) -- parity first, then timings per switch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
python tools/pair_variants.py "PARM_B200_TILE_HALF=0 PARM_B200_TILE_PF=0" "PARM_B200_TILE_HALF=0" "PARM_B200_TILE_HALF=1 PARM_B200_TILE_PF=0" "" 2>gpurun_out/r2s.err | tee gpurun_out/r2s_pair_variants.jsonl
tail -3 gpurun_out/r2s.err
