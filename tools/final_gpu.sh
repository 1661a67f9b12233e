#!/bin/bash
# end-of-iteration evidence: full GPU test suite, smoke, bench line, reference arm, ncu launch list of the bench command
T=${1:-r01_m}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${T}_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/${T}_bench_1M.json 2> gpurun_out/${T}_bench_1M.err; tail -2 gpurun_out/${T}_bench_1M.err; head -c 420 gpurun_out/${T}_bench_1M.json; echo
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_reference.json 2>/dev/null; head -c 300 gpurun_out/${T}_bench_reference.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_bench_1M.csv python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_bench.log 2>&1; tail -1 gpurun_out/${T}_ncu_bench.log | head -c 200; echo
python tools/config_timings.py > gpurun_out/${T}_config_timings.jsonl 2> gpurun_out/${T}_config_timings.err; wc -l gpurun_out/${T}_config_timings.jsonl
python tools/small_n_probe.py > gpurun_out/${T}_small_systems.jsonl 2>> gpurun_out/${T}_config_timings.err; cut -c1-160 gpurun_out/${T}_small_systems.jsonl
python tools/pair_variants.py "" "PARM_B200_TILE_HALF=0 PARM_B200_TILE_PF=0" > gpurun_out/${T}_pair_kernel_variants.jsonl 2>> gpurun_out/${T}_config_timings.err; cut -c1-200 gpurun_out/${T}_pair_kernel_variants.jsonl
# pair-kernel capture + the traffic file bench.py reads (tied to the kernel sources by hash): bash tools/ncu_tile_source.sh ${T}_tile && python tools/ncu_traffic.py ...

