timeout 900 python -m pytest tests/test_gpu_hotpath.py -x -q -m gpu -k "apply" 2>&1 | tail -2
for single in 1 0; do
  echo "single=$single"; ASGFEM_BLK_SINGLE=$single ASGFEM_BLK_VERBOSE=1 timeout 300 python bench.py --variant 9 --steps 5 --warmup 3 --no-e2e --no-cpu --no-pcg --no-est 2> gpurun_out/r2_v9_bench_s$single.err | tail -1 | grep -o '"ms_per_step": [0-9.]*'
  grep "\[blk\]" gpurun_out/r2_v9_bench_s$single.err | tail -1
done
