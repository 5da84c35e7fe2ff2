for f in 0 1; do
if [ $f = 1 ]; then export ASGFEM_BLK_FULLKS=1; fi
ASGFEM_BENCH_NX=128 ASGFEM_BENCH_C5_MINWORLD=1 ASGFEM_BENCH_C5_NX=256 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu --no-pcg --no-est > gpurun_out/r2_c5one.log 2>&1; tail -1 gpurun_out/r2_c5one.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fullks=$f', d['c5']['ms_per_step'])"
done
