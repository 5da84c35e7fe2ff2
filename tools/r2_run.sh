export ASGFEM_MMA_VERBOSE=1
timeout 900 python -m pytest tests/test_gpu_hotpath.py -x -q -k "apply or layout" 2>&1 | tail -40 > gpurun_out/r2_t1.log
tail -30 gpurun_out/r2_t1.log
unset ASGFEM_MMA_VERBOSE
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-pcg --no-est > gpurun_out/r2_b1.log 2>&1
tail -2 gpurun_out/r2_b1.log | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply_mma -s 2 -c 1 -f -o gpurun_out/r2_mma python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-pcg --no-est > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log | cut -c1-300
