set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "spatial or D or golden or wide or shipped or fullsize or generic" 2>&1 | tail -4 > gpurun_out/r02i_tests.log
timeout 300 python bench.py --workload D --steps 128 --warmup 32 --no-cpu > gpurun_out/r02i_bench_D.json 2> gpurun_out/r02i_bench_D.err
tail -3 gpurun_out/r02i_tests.log
