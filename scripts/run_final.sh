# usage: bash scripts/run_final.sh <tag>  -- GPU test suite, smoke(), the driver's bench command (both arms), then the
# captures the bench line's roofline block refers to: launch list of the bench command and --set full of
# step_kernel / reward_kernel (summaries + per-line profiles; the reports themselves are too large to bring back)
set -x
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 128 --warmup 16 --no-cpu --no-secondary > gpurun_out/${TAG}_bench128.json 2> gpurun_out/${TAG}_bench128.err
CMD="python bench.py --steps 8 --warmup 3 --no-cpu --no-secondary"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"step_kernel|reward_half_kernel|reward_kernel" -s 26 -c 4 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep > gpurun_out/ncu_${TAG}_summary.txt 2>&1
for K in step_kernel reward_half_kernel; do
  ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv --print-source cuda,sass -k regex:$K -c 1 > gpurun_out/src_$TAG.csv 2>/dev/null
  python scripts/ncu_lines.py gpurun_out/src_$TAG.csv 60 > gpurun_out/ncu_${TAG}_${K}_lines.txt 2>&1
done
rm -f gpurun_out/src_$TAG.csv gpurun_out/prof_$TAG.ncu-rep
cat gpurun_out/${TAG}_tests.log; tail -2 gpurun_out/${TAG}_smoke.log
python scripts/bench_summary.py gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_bench128.json
