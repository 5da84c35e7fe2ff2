export ASGFEM_BENCH_NX=128 ASGFEM_BENCH_C5_MINWORLD=2 ASGFEM_BENCH_C5_NX=64 ASGFEM_BENCH_C5_N=600
for ord in 2 1; do
ASGFEM_BENCH_C5_ORDER=$ord timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-pcg > gpurun_out/r2_c5small_$ord.log 2>&1
tail -1 gpurun_out/r2_c5small_$ord.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('c5')); print(d['value'], d['config']['sharded_operator_symmetry_defect'])"
done
