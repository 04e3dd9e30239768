#!/usr/bin/env python
"""Experiment: the cfg2 batch as S interleaved sub-batches on S streams (one engine each), so that the tail of one
sub-batch's persistent FFBS kernel is filled by the next kernel of another.  python tools/split_probe.py [S] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from smjp_b200.engine import Engine
from smjp_b200.chains import DeviceChains

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
C, K = bench.CHAINS_PER_GPU, 3
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
obs_t = np.arange(bench.OBS_DT, bench.T_END, bench.OBS_DT)
emis = np.ones((K, K)) + 9 * np.eye(K); emis /= emis.sum(1, keepdims=True)
streams = [torch.cuda.Stream(dev) for _ in range(S)]
parts = []
for s in range(S):
    lo, hi = s * C // S, (s + 1) * C // S
    with torch.cuda.stream(streams[s]):
        eng = Engine(0, stream=streams[s].cuda_stream)
        eng.set_model(bench.SHAPE, np.ones((K, K)), bench.OMEGA, emis, 1.0, bench.T_END)
        eng.set_option("chain_base", lo)
        ch = DeviceChains(eng, hi - lo, obs_t, chain_base=lo)
        ch.init_prior(np.ones(K) / K, seed=bench.SEED + 1).simulate_observations(seed=bench.SEED + 2)
        ch.init_prior(np.ones(K) / K, seed=bench.SEED + 3)
        parts.append((eng, ch))
torch.cuda.synchronize()
for sw in range(3):
    for s, (eng, ch) in enumerate(parts):
        with torch.cuda.stream(streams[s]):
            ch.sweep(seed=bench.SEED, sweep=sw, precision="f64")
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True)
ends = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
e0.record(torch.cuda.current_stream())
for s in range(S):
    streams[s].wait_event(e0)
t0 = time.perf_counter()
ss = 0.0
snaps = []
for i in range(steps):
    for s, (eng, ch) in enumerate(parts):
        with torch.cuda.stream(streams[s]):
            ch.sweep(seed=bench.SEED, sweep=3 + i, precision="f64")
            snaps.append(ch.cur["offs"].clone())
for s in range(S):
    ends[s].record(streams[s])
torch.cuda.synchronize()
ms = max(e0.elapsed_time(e) for e in ends)
for offs in snaps:
    n = (offs[1:] - offs[:-1] - 1).cpu().numpy().astype(np.float64)
    ss += float(np.sum(K * n * (n + 1) / 2))
bad = sum(int((ch.status != 0).sum().item()) for _, ch in parts)
tis = torch.cat([ch.time_in_state for _, ch in parts]).sum().item()
print("S=%d steps=%d: %.3f ms per sweep of %d chains, %.4e state-steps/s, failed %d, time-in-state checksum %.9f" % (S, steps, ms / steps, C, ss / (ms / 1e3), bad, tis))
