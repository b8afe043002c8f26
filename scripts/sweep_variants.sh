# runs bench.py (short) for the default library and every variants/*.so; prints one line each
for lib in default variants/*.so; do
  if [ "$lib" = default ]; then unset PCB_LIB_PATH; else export PCB_LIB_PATH=$PWD/$lib; fi
  out=$(timeout 300 python bench.py --no-cpu --no-secondary --steps 128 --warmup 16 "$@" 2>&1 | tail -n 1)
  echo "$lib $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d["roofline"]
    print("value %.4g  ms/step %.4f  step_us %.1f  alone_us %.1f  frac %.3f whole %.3f e2e %.4g sync %.4g" % (d["value"], d["ms_per_step"], r["avg_launch_us"], r["alone"]["avg_launch_us"], r["frac"], r["whole_step_frac"], d["e2e"]["value"], d["e2e_sync"]["value"]))
except Exception as ex:
    print("FAILED", ex)')"
done
