set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02k_tests.log
for wl in C D E; do
  N=65536; [ $wl = E ] && N=131072
  timeout 300 python bench.py --workload $wl --envs $N --steps 128 --warmup 32 --no-cpu > gpurun_out/r02k_bench_$wl.json 2> gpurun_out/r02k_bench_$wl.err
done
tail -3 gpurun_out/r02k_tests.log
