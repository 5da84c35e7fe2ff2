for w in 16 32; do
  echo "warps=$w"; ASGFEM_BLK_WARPS=$w ASGFEM_BLK_VERBOSE=1 timeout 300 python bench.py --variant 9 --steps 5 --warmup 3 --no-e2e --no-cpu --no-pcg --no-est 2> gpurun_out/r2_v9_bench_$w.err | tail -1 | grep -o '"ms_per_step": [0-9.]*'
  grep "\[blk\]" gpurun_out/r2_v9_bench_$w.err | tail -1
done
