#!/usr/bin/env python
"""the cfg5 block of bench.py alone, one GPU: python tools/cfg5_probe.py [interleave ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
for il in [int(x) for x in sys.argv[1:]] or [1, 2]:
    r = bench.cfg5_block(0, 1, 0, dev, 5, torch.cuda.synchronize, True, interleave=il)
    for p in ("f64", "f32"):
        print("interleave %d %s: %.2f ms per sweep, %.3e state-steps/s, %.3g chain-sweeps/s, failed %d, sha1 %s" % (
            il, p, r[p]["ms_per_step"], r[p]["value"], r[p]["chain_sweeps_per_sec"], r[p]["failed_chains"], r[p]["summaries_sha1"][:10]))
