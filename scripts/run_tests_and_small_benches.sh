set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02f_tests.log
for cfg in "B 4096" "B 65536" "A 65536" "REF_rect 65536"; do
  set -- $cfg
  timeout 300 python bench.py --workload $1 --envs $2 --steps 256 --warmup 32 --no-cpu > gpurun_out/r02f_bench_$1_$2_fused.json 2> gpurun_out/r02f_bench_$1_$2_fused.err
done
tail -3 gpurun_out/r02f_tests.log
