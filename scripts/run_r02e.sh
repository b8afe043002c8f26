set -x
mkdir -p gpurun_out
for cfg in "B 4096" "B 65536" "A 65536" "REF_rect 65536"; do
  set -- $cfg
  timeout 300 python bench.py --workload $1 --envs $2 --steps 256 --warmup 32 --no-cpu > gpurun_out/r02e_bench_$1_$2_fused.json 2> gpurun_out/r02e_bench_$1_$2_fused.err
done
bash scripts/prof_small.sh r02e B 65536
