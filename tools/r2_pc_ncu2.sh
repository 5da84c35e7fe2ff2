NX=1024 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_small_step -c 5 -f -o gpurun_out/r2_trsv24 python tools/prof_trsv.py > gpurun_out/r2_pc_ncu2.log 2>&1
tail -2 gpurun_out/r2_pc_ncu2.log
