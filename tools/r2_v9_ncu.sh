ASGFEM_BLK_WARPS=${W:-16} timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply_blk -s 2 -c 1 -f -o gpurun_out/r2_v9 python bench.py --variant 9 --steps 1 --warmup 3 --no-e2e --no-cpu --no-pcg --no-est > gpurun_out/r2_v9_ncu.log 2>&1
tail -2 gpurun_out/r2_v9_ncu.log | cut -c1-300
