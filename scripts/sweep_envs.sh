# env-steps/s against the number of envs per GPU (device-timed value and e2e), configs C and B
mkdir -p gpurun_out
for N in 1024 4096 16384 65536 131072 262144; do
  python bench.py --workload C --envs $N --steps 128 --warmup 16 --no-cpu --no-secondary > gpurun_out/sweepN_C_$N.json 2>/dev/null
done
for N in 1024 4096 16384 65536 262144 1048576; do
  python bench.py --workload B --envs $N --steps 128 --warmup 16 --no-cpu > gpurun_out/sweepN_B_$N.json 2>/dev/null
done
