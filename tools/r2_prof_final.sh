# launch list of the bench command (per-launch gpu time; cold caches, serialised) and a full capture of the default operator kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-pcg --no-est > gpurun_out/r02_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply_ts2 -s 2 -c 1 -f -o gpurun_out/r02_apply_ts2 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-pcg --no-est > gpurun_out/r02_apply_ts2_ncu.log 2>&1
tail -1 gpurun_out/r02_launches_bench.log | cut -c1-200
