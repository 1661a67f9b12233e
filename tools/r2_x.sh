#!/bin/bash
# 2 GPUs: tests/mgpu_check.py including the migration-path equivalence check (one sort / overflow fall-back / two sorts)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | grep -v Warning | grep "mgpu ok\|MGPU\|Error\|error\|assert\|Traceback" | tee gpurun_out/r2x_mgpu_check_2gpu.log | cut -c1-600
