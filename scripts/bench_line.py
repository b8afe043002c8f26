import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line); r=d["roofline"]
    print("%-60s value %.4g  ms/step %.4f  step_us %.1f  frac %.3f whole %.3f e2e %.4g" % (d["config"]["workload"][:60]+" N=%d"%d["config"]["envs_per_gpu"], d["value"], d["ms_per_step"], r["avg_launch_us"], r["frac"], r["whole_step_frac"], d["e2e"]["value"]))
