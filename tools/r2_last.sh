timeout 200 python -m pytest tests/test_gpu_hotpath.py -x -q -m gpu -k "apply" 2>&1 | tail -2
