# usage: bash scripts/prof_step.sh <tag> [extra bench args]
set -x
TAG=$1; shift
CMD="python bench.py --steps 8 --warmup 3 --no-cpu $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"step_kernel" -s 14 -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -n 3 gpurun_out/ncu_$TAG.log
