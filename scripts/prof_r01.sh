set -x
CMD="python bench.py --steps 8 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"step_kernel|terminal_kernel|sample_kernel" -s 42 -c 6 -o gpurun_out/prof_r01a $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/plain.log gpurun_out/ncu1.log gpurun_out/ncu2.log
ls -la gpurun_out
