#!/usr/bin/env python
"""one-line digest of a bench.py JSON line on stdin (kernel experiments)"""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d["roofline"]
print("rescue %s " % d.get("rescue"), end="")
print("ms/step %.3f  kernel %s %.3f ms  grid %s threads %s smem %s  frac %.4f live %.4f  failed %s" % (
    d["ms_per_step"], r["kernel"], r["kernel_ms_avg"], r.get("grid"), r.get("threads"), r.get("smem"), r["frac"], r["live_fraction"], d["failed_chains"]))
