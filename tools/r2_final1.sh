timeout 1500 python bench.py > gpurun_out/r2_final_bench1.log 2> gpurun_out/r2_final_bench1.err; tail -1 gpurun_out/r2_final_bench1.log | cut -c1-6000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_ref.log 2>&1; tail -1 gpurun_out/r2_final_ref.log | cut -c1-1500
