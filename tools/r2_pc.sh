timeout 900 python -m pytest tests/test_gpu_hotpath.py -x -q -m gpu -k "precond or pcg or solve or logprimal or bicg" > gpurun_out/r2_pc_tests.log 2>&1; tail -4 gpurun_out/r2_pc_tests.log
ASGFEM_SWEEP_SPLIT=3 timeout 900 python -m pytest tests/test_gpu_hotpath.py -x -q -m gpu -k "precond" > gpurun_out/r2_pc_tests3.log 2>&1; tail -2 gpurun_out/r2_pc_tests3.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-est > gpurun_out/r2_pc_bench.log 2>&1
tail -1 gpurun_out/r2_pc_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); print(d.get('pcg'))"
