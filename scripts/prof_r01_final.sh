# ncu evidence for profiles/: launch list of the bench command + full capture of the
# step and reward kernels.  Usage: bash scripts/prof_r01_final.sh <tag>
set -x
TAG=$1
CMD="python bench.py --steps 8 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"step_kernel|reward_kernel" -s 28 -c 4 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_f_$TAG.log | cut -c1-200
