timeout 900 python -m pytest tests/test_gpu_hotpath.py -x -q -m gpu -k "logprimal_estimator or estimator_matches" > gpurun_out/r2_f2.log 2>&1; tail -30 gpurun_out/r2_f2.log
