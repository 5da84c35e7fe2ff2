# what the driver runs at round end, on one GPU: GPU tests, smoke, default bench, reference arm
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_tests.log 2>&1; tail -3 gpurun_out/r2_final_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r2_final_bench1.log 2> gpurun_out/r2_final_bench1.err; tail -1 gpurun_out/r2_final_bench1.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'pcg',d['pcg']['solve_s'],d['pcg']['iterations'],'e2e_solve',d['pcg']['e2e_solve']['seconds'],'est',d['estimator']['kernel_ms'],d['estimator']['call_ms'],d['estimator']['call_ms_marking_outputs'],'log',d['logprimal']['ms_per_iteration'],'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['value_1thread'])"
