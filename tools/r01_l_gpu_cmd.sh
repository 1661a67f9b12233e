set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r01_l_bench_1M.json 2> gpurun_out/r01_l_bench_1M.err; tail -c 1500 gpurun_out/r01_l_bench_1M.json
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "probes" 2>&1 | tail -3
python bench.py --side 200 --side-z 400 --steps 50 --warmup 5 --e2e-steps 3 --no-cpu > gpurun_out/r01_l_bench_16M_1gpu.json 2> gpurun_out/r01_l_bench_16M_1gpu.err; tail -c 600 gpurun_out/r01_l_bench_16M_1gpu.err; head -c 400 gpurun_out/r01_l_bench_16M_1gpu.json
