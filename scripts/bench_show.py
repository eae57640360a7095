"""Prints the essentials of a bench.py JSON line (helper for reading gpurun_out/*.json)."""
import json
import sys

for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e:
        print(p, "unreadable:", e)
        continue
    print(p)
    print("  value %.5f s  e2e %.5f s  launches %s  iterations %s  setup %.1f s" % (
        d["value"], d["e2e"]["value"], d.get("gpu_launches"), d["config"].get("iterations"), d["config"].get("setup_seconds_untimed", 0)))
    print("  ", d["config"].get("l2"), d.get("clocks"))
    for k, v in d.get("operators", {}).items():
        print("   %-28s %9.4f ms %7.0f GB/s  frac %.3f" % (k, v["ms"], v["gbs"], v["frac_of_peak"]))
    print("  ", d.get("time_share_seconds_profiled_solve"))
    if "cpu_baseline" in d:
        print("  cpu:", d["cpu_baseline"])
