timeout 900 python -m pytest tests/test_gpu_hotpath.py -x -q -m gpu -k "deterministic_sample or evaluate_samples" > gpurun_out/r2_f4.log 2>&1; tail -25 gpurun_out/r2_f4.log
