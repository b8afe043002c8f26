# per-step kernels (default above 12 288 envs) against the multi-step pin kernel forced (PCB_PIN_FUSED_MAX_ENVS) on C / D
for W in C D; do
for N in 16384 32768 65536; do
  for mode in default fused; do
    if [ $mode = fused ]; then export PCB_PIN_FUSED_MAX_ENVS=100000000; else unset PCB_PIN_FUSED_MAX_ENVS; fi
    out=$(timeout 300 python bench.py --no-cpu --no-secondary --steps 128 --warmup 16 --workload $W --envs $N 2>&1 | tail -n 1)
    echo "$mode $(echo "$out" | python scripts/bench_line.py)"
  done
done
done
