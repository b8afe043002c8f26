# round-2 final captures: launch list of the bench command + --set full of the main kernels
# (run under gpurun; ncu numbers are cold-cache and serialised -- compare shares, not absolutes).
# gpurun brings back at most 64 MiB: every report is summarised on the box (scripts/ncu_summary.py,
# per-line profile of the hot kernel) and only the main report is kept.
set -x
TAG=$1
mkdir -p gpurun_out
summarise() {  # <report tag> <kernel regex for the per-line profile>
  python scripts/ncu_summary.py gpurun_out/prof_$1.ncu-rep > gpurun_out/ncu_${1}_summary.txt 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page source --csv --print-source cuda,sass -k regex:"$2" -c 1 > gpurun_out/src_$1.csv 2>/dev/null
  python scripts/ncu_lines.py gpurun_out/src_$1.csv 40 > gpurun_out/ncu_${1}_lines.txt 2>&1
  rm -f gpurun_out/src_$1.csv
}
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-secondary"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"step_kernel|reward_kernel" -s 26 -c 4 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
summarise $TAG step_kernel
for W in D E; do
  C2="python bench.py --steps 8 --warmup 3 --no-cpu --no-secondary --workload $W"
  $C2 > gpurun_out/plain_${TAG}_$W.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"step_kernel" -s 40 -c 1 -o gpurun_out/prof_${TAG}_$W $C2 > gpurun_out/ncu_f_${TAG}_$W.log 2>&1
  summarise ${TAG}_$W step_kernel; rm -f gpurun_out/prof_${TAG}_$W.ncu-rep
done
C3="python bench.py --steps 64 --warmup 3 --no-cpu --no-secondary --workload B --envs 65536"
$C3 > gpurun_out/plain_${TAG}_B.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_small -s 7 -c 1 -o gpurun_out/prof_${TAG}_B $C3 > gpurun_out/ncu_f_${TAG}_B.log 2>&1
summarise ${TAG}_B rollout_small; rm -f gpurun_out/prof_${TAG}_B.ncu-rep
C4="python bench.py --steps 64 --warmup 3 --no-cpu --no-secondary --workload C --envs 4096"
$C4 > gpurun_out/plain_${TAG}_C4096.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rollout_pin_kernel|reward_list" -s 22 -c 2 -o gpurun_out/prof_${TAG}_C4096 $C4 > gpurun_out/ncu_f_${TAG}_C4096.log 2>&1
summarise ${TAG}_C4096 rollout_pin_kernel; rm -f gpurun_out/prof_${TAG}_C4096.ncu-rep
tail -n 2 gpurun_out/ncu_f_$TAG.log
ls -la gpurun_out
