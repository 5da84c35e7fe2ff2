timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu -k estimator > gpurun_out/r2_est2.log 2>&1; tail -25 gpurun_out/r2_est2.log
