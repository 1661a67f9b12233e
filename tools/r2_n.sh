#!/bin/bash
# batched stepping: parity, then all single-GPU configurations with and without
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py tests/test_gpu_scale.py tests/test_gpu_edge.py tests/test_gpu_golden.py tests/test_gpu_reference_tests.py tests/test_gpu_pyparm.py -m gpu -x -q 2>&1 | tail -5
for b in 0 16; do
  echo "== PARM_B200_STEP_BATCH=$b"
  PARM_B200_STEP_BATCH=$b python tools/config_timings.py --steps 400 2>gpurun_out/r2n.err | tee gpurun_out/r2n_configs_batch$b.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%-40s N %8d  %.4g atom-steps/s  ms/step %.5f  k1 %.4f force %.4f k3 %.4f rebuild %.3f x%d' % (d['config'][:40], d['n_atoms'], d['atom_steps_per_s'], d['ms_per_step'], d['k1_ms'], d['force_ms'], d['k3_ms'], d['rebuild_ms_each'], d['rebuilds']))
"
  tail -2 gpurun_out/r2n.err
done
