timeout 900 python -m pytest tests/test_gpu_hotpath.py -x -q -m gpu -k "estimator" 2>&1 | tail -3
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-pcg > gpurun_out/r2_est_bench.log 2>&1; tail -1 gpurun_out/r2_est_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('estimator'))"
