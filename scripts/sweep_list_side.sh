# list_steps producer (reward types beam / both, large batches): side-grid size of the router / reward list kernel
for W in C_both C_beam; do
  echo "## $W"
  for M in 1 2 4 8 16; do
    out=$(PCB_LIST_STEPS_MIN_ENVS=1 PCB_LIST_SIDE_MULT=$M timeout 300 python bench.py --no-cpu --no-secondary --steps 128 --warmup 16 --workload $W 2>&1 | tail -n 1)
    echo "side_mult $M $(echo "$out" | python scripts/bench_line.py)"
  done
  out=$(PCB_LIST_STEPS_MIN_ENVS=100000000 timeout 300 python bench.py --no-cpu --no-secondary --steps 128 --warmup 16 --workload $W 2>&1 | tail -n 1)
  echo "rollout_pin_kernel $(echo "$out" | python scripts/bench_line.py)"
done
