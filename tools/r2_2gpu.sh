timeout 500 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -25 > gpurun_out/r2_dist.log
tail -25 gpurun_out/r2_dist.log
