timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_tests.log 2>&1; tail -5 gpurun_out/r2_tests.log
