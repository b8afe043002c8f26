# usage: bash scripts/run_full.sh <tag>  -- whole GPU test suite + the bench lines of every workload
set -x
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --cpu-seconds 5 > gpurun_out/${TAG}_bench_C.json 2> gpurun_out/${TAG}_bench_C.err
for wl in C_both C_beam E REF_pin; do
  N=65536; [ $wl = E ] && N=131072
  timeout 300 python bench.py --workload $wl --envs $N --steps 128 --warmup 32 --no-cpu > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
done
tail -3 gpurun_out/${TAG}_tests.log
