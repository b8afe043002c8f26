# final captures of a round: launch list of the bench command + --set full of the three main kernels
set -x
TAG=$1
mkdir -p gpurun_out
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-secondary"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"step_kernel|reward_kernel" -s 26 -c 4 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
bash scripts/prof_small.sh ${TAG}_B B 65536
tail -n 2 gpurun_out/ncu_f_$TAG.log
