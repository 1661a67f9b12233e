#!/bin/bash
# small systems: gather kernel with 16 lanes per atom (default for < 148 chunks) against the previous choice
mkdir -p gpurun_out
for v in "PARM_B200_SMALL_TEAM=0" ""; do
  env $v python tools/small_n_probe.py 2>>gpurun_out/r2v.err | tee -a gpurun_out/r2v_small_n.jsonl
done
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity.py tests/test_gpu_reference_tests.py -m gpu -x -q 2>&1 | tail -3
tail -3 gpurun_out/r2v.err
