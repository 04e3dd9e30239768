#!/usr/bin/env python
"""K = 10 kernel experiments: python tools/k10_probe.py key=value ... (engine options), prints the k10 block of bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
opts = {}
chains = bench.K10_CHAINS
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k == "chains":
        chains = int(v)
    else:
        opts[k] = int(v)
peak, src = bench.measured_peak()
r = bench.k10_block(0, 4, peak, src, chains=chains, opts=opts)
for p in ("f64", "f32"):
    x = r[p]
    print(opts, p, "ms/step %.2f kernel %.2f ms %s frac %.3f live %.3f nominal_frac %.3f rescued %d failed %d grid %d smem %d" % (
        x["ms_per_step"], x["roofline"]["kernel_ms_avg"], x["roofline"]["kernel"], x["roofline"]["frac"], x["roofline"]["live_fraction"],
        x["roofline"]["nominal_frac"], x["chains_redone_by_rescue_pass_last_launch"], x["failed_chains"], x["roofline"]["grid"], x["roofline"]["smem"]))
