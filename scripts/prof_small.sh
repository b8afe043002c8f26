# usage: bash scripts/prof_small.sh <tag> <workload> <envs>   -- ncu --set full of the timed rollout_small_kernel launch
set -x
mkdir -p gpurun_out
TAG=$1; WL=$2; N=$3
CMD="python bench.py --workload $WL --envs $N --steps 64 --warmup 3 --no-cpu"
SKIP=$(python -c "import bench; print(bench.EP_LEN['$WL'] + 1)")
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_small -s $SKIP -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_$TAG.log
