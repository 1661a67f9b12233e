#!/bin/bash
# quick GPU iteration: tile-kernel parity tests, then the gather-vs-tile timing of the bench workload
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -4
python tools/tile_sweep.py --quick 2>gpurun_out/quick.err | tee gpurun_out/quick.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['env'], 'ms/step %.4f force %.4f rebuild %.4f k1 %.4f k3 %.4f E %.10g' % (d['ms_per_step'], d['force_ms'], d['rebuild_ms_each'], d['k1_ms'], d['k3_ms'], d['energy']))
"
tail -3 gpurun_out/quick.err
