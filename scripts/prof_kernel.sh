# usage: bash scripts/prof_kernel.sh <tag> <kernel-regex> <skip> [bench args]
set -x
TAG=$1; KRE=$2; SKIP=$3; shift 3
CMD="python bench.py --steps 8 --warmup 3 --no-cpu $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 2 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_$TAG.log
