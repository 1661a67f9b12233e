timeout 900 python -m pytest tests/test_gpu_pyparm.py tests/test_gpu_facade.py -m gpu -x -q 2>&1 | tail -15
