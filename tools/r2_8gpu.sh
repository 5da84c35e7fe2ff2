timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench8.log 2> gpurun_out/r2_bench8.err
tail -1 gpurun_out/r2_bench8.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['sharded_operator_symmetry_defect']); print(d.get('pcg')); print(d.get('c5'))"
tail -3 gpurun_out/r2_bench8.err
