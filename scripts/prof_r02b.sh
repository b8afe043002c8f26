set -x
mkdir -p gpurun_out
for N in 4096 65536; do
CMD="python bench.py --workload B --envs $N --steps 64 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_r02b_$N.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_small -s 7 -c 1 -o gpurun_out/prof_r02b_$N $CMD > gpurun_out/ncu_r02b_$N.log 2>&1
tail -n 2 gpurun_out/ncu_r02b_$N.log
done
