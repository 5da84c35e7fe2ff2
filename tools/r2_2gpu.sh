timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | tail -15 > gpurun_out/r2_dist.log
tail -15 gpurun_out/r2_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench2.log 2>&1
tail -3 gpurun_out/r2_bench2.log | cut -c1-3000
