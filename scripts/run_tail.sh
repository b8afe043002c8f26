# usage: bash scripts/run_tail.sh <tag> -- parity suite, then the committed library against variants/libpcbgpu_prev.so:
# rows-per-call fit of pcb_steps_host and the driver-style bench line (--steps 20 --warmup 5)
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${TAG}_tests.log
for lib in default variants/libpcbgpu_prev.so; do
  if [ "$lib" = default ]; then unset PCB_LIB_PATH; else export PCB_LIB_PATH=$PWD/$lib; fi
  echo "## $lib" >> gpurun_out/${TAG}_rows.log
  python scripts/e2e_rows_per_call.py >> gpurun_out/${TAG}_rows.log 2>&1
done
for r in 1 2; do
for lib in default variants/libpcbgpu_prev.so; do
  if [ "$lib" = default ]; then unset PCB_LIB_PATH; else export PCB_LIB_PATH=$PWD/$lib; fi
  echo "## $lib" >> gpurun_out/${TAG}_bench20.log
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary 2>/dev/null | tail -n 1 | python scripts/bench_line.py >> gpurun_out/${TAG}_bench20.log 2>&1
done
done
cat gpurun_out/${TAG}_tests.log gpurun_out/${TAG}_rows.log gpurun_out/${TAG}_bench20.log
