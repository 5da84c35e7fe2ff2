for s in 0 1 2 3 4 7; do
  echo "skip=$s"; ASGFEM_MMA_SKIP=$s timeout 300 python bench.py --variant 8 --steps 3 --warmup 3 --no-e2e --no-cpu --no-pcg --no-est 2>/dev/null | tail -1 | grep -o '"ms_per_step": [0-9.]*'
done
