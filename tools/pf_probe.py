#!/usr/bin/env python
"""cfg3 particle-filter block of bench.py alone: python tools/pf_probe.py [family] [runs ...] [option=value ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
family = sys.argv[1] if len(sys.argv) > 1 else "gamma"
opts = {a.split("=")[0]: int(a.split("=")[1]) for a in sys.argv[2:] if "=" in a}
runs = [int(x) for x in sys.argv[2:] if "=" not in x] or [8, 128]
peak, src = bench.measured_peak()
dev = torch.device("cuda", 0)
for r in runs:
    x = bench.pf_block(0, 1, 0, dev, r, peak, src, family=family, opts=opts)
    print(opts, family, "runs %d: %.1f ms/call  %.3g particle-steps/s  frac %.4f  loglik %.4f +- %.4f  failed %d cluster %s" % (
        r, x["ms_per_call"], x["value"], x["roofline"]["frac"], x["loglik_mean"], x["loglik_std_over_runs"], x["failed_runs"], x["cluster_size"]))
