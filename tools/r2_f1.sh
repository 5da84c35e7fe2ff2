timeout 900 python -m pytest tests/test_gpu_hotpath.py -x -q -m gpu -k "logprimal" > gpurun_out/r2_f1.log 2>&1; tail -30 gpurun_out/r2_f1.log
