ASGFEM_BENCH_NX=128 ASGFEM_BENCH_C5_MINWORLD=1 ASGFEM_BENCH_C5_NX=128 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply_blk -s 5 -c 1 -f -o gpurun_out/r2_c5 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu --no-pcg --no-est > gpurun_out/r2_c5ncu.log 2>&1
tail -2 gpurun_out/r2_c5ncu.log | cut -c1-200
