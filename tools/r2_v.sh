#!/bin/bash
# small systems: persistent cooperative kernel (small.cu) against the per-step kernels
mkdir -p gpurun_out
for v in "" "PARM_B200_SMALL_PERSIST=0"; do
  env $v timeout 180 python tools/small_n_probe.py 2>>gpurun_out/r2v.err | tee -a gpurun_out/r2v_small_n.jsonl | cut -c1-330
done
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_reference_tests.py tests/test_gpu_facade.py tests/test_gpu_pyparm.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -5
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
tail -3 gpurun_out/r2v.err
