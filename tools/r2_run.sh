timeout 900 python -m pytest tests/test_gpu_hotpath.py -x -q -k "apply or layout" 2>&1 | tail -30 > gpurun_out/r2_t1.log
tail -30 gpurun_out/r2_t1.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-pcg --no-est > gpurun_out/r2_b1.log 2>&1
tail -2 gpurun_out/r2_b1.log | cut -c1-400
bash tools/r2_skip.sh
