# variant 9 (block kernel with list exchange): parity tests + bench, 16 and 32 warps
timeout 900 python -m pytest tests/test_gpu_hotpath.py -x -q -m gpu -k "apply" > gpurun_out/r2_v9_tests.log 2>&1; tail -5 gpurun_out/r2_v9_tests.log
for w in 16; do
  echo "warps=$w"; ASGFEM_BLK_WARPS=$w ASGFEM_BLK_VERBOSE=1 timeout 300 python bench.py --variant 9 --steps 5 --warmup 3 --no-e2e --no-cpu --no-pcg --no-est 2> gpurun_out/r2_v9_bench_$w.err | tail -1 | grep -o '"ms_per_step": [0-9.]*'
  grep "\[blk\]" gpurun_out/r2_v9_bench_$w.err | tail -3
done
