"""One line per bench JSON file: value, ms/step, e2e, roofline fractions."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(path.split("/")[-1], "value %.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"],
              "e2e %.3e" % d["e2e"]["value"], "frac %.3f whole %.3f" % (r["frac"], r["whole_step_frac"]),
              "launches", d["gpu_launches"], "sm_mhz", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:  # noqa: BLE001
        print(path, "ERR", e)
