#!/usr/bin/env python
"""Pretty-print bench.py JSON lines: headline, e2e, cpu baseline and the per-kernel table."""
import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        print("== %s: %s  N=%d  %.1f steps/s  %.4f ms/step  S=%s  launches=%s" % (f, d["config"]["workload"], d["n_gpus"], d["value"], d["ms_per_step"], d["config"].get("n_states"), d.get("gpu_launches")))
        print("   e2e:", d.get("e2e") and round(d["e2e"]["value"], 1), " cpu:", d.get("cpu_baseline") and (round(d["cpu_baseline"]["value"], 3), d["cpu_baseline"]["cores"]), " clocks:", d.get("clocks"))
        r = d.get("roofline")
        if r: print("   roofline:", r["kernel"], r["bound"], "%.1f %s frac %.3f traffic %s alg %s" % (r["achieved"], r["unit"], r["frac"], r["traffic"], r["algorithmic"]))
        for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
            print("   %-22s %.4f ms  x%.1f  share %.3f  %s" % (k, v["ms_per_step"], v["launches_per_step"], v["share_of_step"] or 0, ("%s %.4g %s frac %.3f" % (v["bound"], v["achieved"], v["unit"], v["frac"])) if "bound" in v else ""))
