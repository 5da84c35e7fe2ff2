#!/bin/bash
# tools/run_2gpu_grid.sh: 2-GPU bench with different persistent-grid sizes of the operator kernel
for g in 140 148; do
  export ASGFEM_TS2_GRID=$g
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('grid', os.environ['ASGFEM_TS2_GRID'], d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
done
