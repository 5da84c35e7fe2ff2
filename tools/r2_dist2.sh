ASGFEM_SWEEP_SPLIT=3 timeout 900 python -m pytest tests/test_gpu_hotpath.py -x -q -m gpu -k "precond or pcg or solve" > gpurun_out/r2_split_tests.log 2>&1; tail -3 gpurun_out/r2_split_tests.log
timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu > gpurun_out/r2_dist.log 2>&1; tail -3 gpurun_out/r2_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench2.log 2>gpurun_out/r2_bench2.err
tail -1 gpurun_out/r2_bench2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['sharded_operator_symmetry_defect']); print(d.get('pcg'))"
