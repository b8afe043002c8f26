# usage: bash scripts/run_half.sh <tag> -- parity suite, then default (two jobs per warp) against variants/*.so
# (g8 = four jobs per warp, half0 = one) on C, D, C at 4 096 envs and the shipped 10x10 pin config,
# then --set full + per-line profile of reward_half_kernel
set -x
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_tests.log
for r in 1 2; do
bash scripts/sweep_variants.sh >> gpurun_out/${TAG}_sweep_C.log 2>&1
done
bash scripts/sweep_variants.sh --workload D > gpurun_out/${TAG}_sweep_D.log 2>&1
bash scripts/sweep_variants.sh --workload REF_pin --envs 65536 > gpurun_out/${TAG}_sweep_REF_pin.log 2>&1
bash scripts/sweep_variants.sh --workload C --envs 4096 > gpurun_out/${TAG}_sweep_C4096.log 2>&1
unset PCB_LIB_PATH
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-secondary"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"reward_half" -s 13 -c 1 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_f_${TAG}.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_${TAG}.ncu-rep > gpurun_out/ncu_${TAG}_reward_summary.txt 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv --print-source cuda,sass -k regex:reward_half -c 1 > gpurun_out/src_${TAG}.csv 2>/dev/null
python scripts/ncu_lines.py gpurun_out/src_${TAG}.csv 60 > gpurun_out/ncu_${TAG}_reward_lines.txt 2>&1
rm -f gpurun_out/src_${TAG}.csv gpurun_out/prof_${TAG}.ncu-rep
cat gpurun_out/${TAG}_tests.log gpurun_out/${TAG}_sweep_*.log
