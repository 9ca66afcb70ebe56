"""one-line digest of a bench.py JSON line read from stdin (used by the sweep scripts)"""
import json
import sys

d = json.loads(sys.stdin.readline())
r = d["roofline"]
vol = {k: v for k, v in d.get("refine_volumes_last_step", {}).items() if k != "refine_phase_cycles"}
print("%.0f q/s  %.3f ms/step  e2e %.3f ms  waves %d  reruns %d (reason %s)  scan %.3f ms  %s %.0f %s frac %.3f  vol %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["waves_per_step"], d["overflow_reruns"],
    d.get("overflow_reason"), r["kernel_ms_per_step"], r["bound"], r["achieved"], r["unit"], r["frac"], vol))
