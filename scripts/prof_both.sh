# usage: bash scripts/prof_both.sh <tag> -- --set full of the side kernels of the multi-step pin path with reward type
# "both" (config C, 65 536 envs): beam_kernel and reward_list_half_kernel<8, true> over the job list of an 18-step launch,
# and of rollout_pin_kernel itself; summaries only (the reports are too large to bring back)
set -x
TAG=$1
CMD="python bench.py --steps 36 --warmup 3 --no-cpu --no-secondary --workload C_both"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"beam_kernel|reward_list_half_kernel|rollout_pin_kernel" -s 36 -c 3 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep > gpurun_out/ncu_${TAG}_summary.txt 2>&1
for K in beam_kernel reward_list_half_kernel; do
  ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv --print-source cuda,sass -k regex:$K -c 1 > gpurun_out/src_$TAG.csv 2>/dev/null
  python scripts/ncu_lines.py gpurun_out/src_$TAG.csv 40 > gpurun_out/ncu_${TAG}_${K}_lines.txt 2>&1
done
rm -f gpurun_out/src_$TAG.csv gpurun_out/prof_$TAG.ncu-rep
tail -n 2 gpurun_out/ncu_f_$TAG.log
grep -E "Kernel Name|time_duration|inst_executed.sum|registers|issue_active|grid_size" gpurun_out/ncu_${TAG}_summary.txt
