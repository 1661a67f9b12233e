python -m pytest tests/test_gpu_edge.py -m gpu -x -q -k row_capacity 2>&1 | grep -E "ParmError|Error:|error" | head -5
