# usage: bash scripts/run_cross.sh <tag>  -- parity suite, then default library against variants/*.so on C, D, E,
# then a per-line ncu profile of reward_kernel
set -x
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_tests.log
bash scripts/sweep_variants.sh > gpurun_out/${TAG}_sweep_C.log 2>&1
bash scripts/sweep_variants.sh --workload D > gpurun_out/${TAG}_sweep_D.log 2>&1
bash scripts/sweep_variants.sh --workload E --envs 131072 > gpurun_out/${TAG}_sweep_E.log 2>&1
bash scripts/sweep_variants.sh --workload C_both > gpurun_out/${TAG}_sweep_C_both.log 2>&1
unset PCB_LIB_PATH
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary > gpurun_out/${TAG}_bench20.json 2> gpurun_out/${TAG}_bench20.err
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-secondary"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"reward_kernel" -s 13 -c 1 -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_f_${TAG}.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_${TAG}.ncu-rep > gpurun_out/ncu_${TAG}_reward_summary.txt 2>&1
ncu -i gpurun_out/prof_${TAG}.ncu-rep --page source --csv --print-source cuda,sass -k regex:reward_kernel -c 1 > gpurun_out/src_${TAG}.csv 2>/dev/null
python scripts/ncu_lines.py gpurun_out/src_${TAG}.csv 60 > gpurun_out/ncu_${TAG}_reward_lines.txt 2>&1
rm -f gpurun_out/src_${TAG}.csv gpurun_out/prof_${TAG}.ncu-rep
cat gpurun_out/${TAG}_tests.log gpurun_out/${TAG}_sweep_*.log
