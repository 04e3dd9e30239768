#!/usr/bin/env python
"""Runs the BASELINE.json configurations that are not the bench line (configs[0], [2], [3], [4]) at
representative sizes on one GPU and prints one JSON line each: they are parity-test configurations
(tests/ cover their correctness); this script records what they cost on a B200.

    python tools/run_configs.py [cfg1 cfg3 cfg4 cfg5]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def emission(K):
    E = np.ones((K, K)) + 9 * np.eye(K)
    return E / E.sum(axis=1, keepdims=True)


def cfg1():
    """main.py default: single 3-state chain, T=2, 19 observations -- through the drop-in raoteh()"""
    import experiments
    from smjp_b200.raoteh import raoteh
    m = experiments.build_experiment_2_model()
    raoteh(["trajectory"], 20, 300, m["state_space"], m["time_final"], m["emission"], m["data"], m["pi_0"],
           m["hazard_A"], m["hazard_B"], m["poisson_process_A_hat"], m["poisson_process_B"], "cfg1", m["omega"],
           m["obs_times"], None, False, seed=0, save=False)  # context creation, module load: not part of a sweep
    n = 2000
    t0 = time.perf_counter()
    agg, _, _ = raoteh(["trajectory"], n, 300, m["state_space"], m["time_final"], m["emission"], m["data"], m["pi_0"],
                       m["hazard_A"], m["hazard_B"], m["poisson_process_A_hat"], m["poisson_process_B"], "cfg1", m["omega"],
                       m["obs_times"], None, False, seed=1, save=False)
    dt = time.perf_counter() - t0
    ss = sum(3 * (len(w) - 1) * len(w) / 2 for w in agg["W"])
    return {"config": "cfg1 main.py default, 1 chain through the drop-in raoteh()", "sweeps_per_sec": n / dt,
            "state_steps_per_sec": ss / dt, "reference_measured_sweeps_per_sec": 1.04, "note": "latency-bound: one fused host call per sweep (smjp_sweep_host)"}


def cfg3(N=65536, C=8, cluster=0, threads=0, family="gamma"):
    """pmcmc particle filter: 65536 particles per chain, 5-state sMJP, systematic resampling"""
    import torch
    from smjp_b200.engine import Engine, pf_host
    rng = np.random.default_rng(3)
    K, t_end = 5, 10.0
    eng = Engine(0)
    eng.set_model(rng.uniform(0.6, 3.0, (K, K)), np.ones((K, K)), 2.0, emission(K), 1.0, t_end)
    obs_t = np.arange(0.1, t_end, 0.1)
    y = rng.integers(0, K, len(obs_t))
    eng.set_option("pf_cluster", cluster)
    eng.set_option("pf_threads", threads)
    pf_host(eng, obs_t, y, np.ones(K) / K, N, seed=1, C=C, family=family)
    torch.cuda.synchronize()
    eng.set_option("time_kernels", 1)
    t0 = time.perf_counter()
    r = pf_host(eng, obs_t, y, np.ones(K) / K, N, seed=2, C=C, family=family)
    dt = time.perf_counter() - t0
    kms, _ = eng.kernel_time("pf")
    cs = eng.lib.smjp_ctx_get_stat(eng.ctx, b"pf_cluster")
    eng.close()
    return {"config": "cfg3 particle filter N=%d, K=5 %s hazards, D=%d, %d filter runs, systematic" % (N, family, len(obs_t), C),
            "cluster": cs, "threads": threads,
            "particle_steps_per_sec": N * len(obs_t) * C / dt, "seconds": dt, "kernel_ms": kms,
            "kernel_particle_steps_per_sec": N * len(obs_t) * C / (kms / 1e3), "loglik_spread": float(np.ptp(r["loglik"])),
            "algorithmic_GBps": 68.0 * N * len(obs_t) * C / dt / 1e9, "status_ok": bool(np.all(r["status"] == 0))}


def cfg4(C=1024, K=64, T=2000.0, sweeps=3):
    """large-state dense path: K=64 exponential sMJP (shape == 1 collapse), Rao-Teh sweeps on the device"""
    import torch
    from smjp_b200.chains import DeviceChains
    from smjp_b200.engine import Engine
    rng = np.random.default_rng(4)
    eng = Engine(0)
    # total leaving rate A_v ~ 1 per unit time: about 2*T grid points per chain at omega = 2
    eng.set_model(np.ones((K, K)), rng.uniform(0.5 * K, 1.5 * K, (K, K)), 2.0, emission(K), 1.0, T)
    obs_t = np.arange(0.5, T, 1.0)
    out = {}
    for prec in ("f64", "f32"):
        ch = DeviceChains(eng, C, obs_t).init_prior(np.ones(K) / K, seed=1, honour_scale=True, cap_per_chain=4096).simulate_observations(seed=2)
        ch.init_prior(np.ones(K) / K, seed=3, honour_scale=True, cap_per_chain=4096)
        for sw in range(2):
            ch.sweep(seed=4, sweep=sw, precision=prec, dense=True)
        torch.cuda.synchronize()
        eng.set_option("time_kernels", 1)
        t0 = time.perf_counter()
        for sw in range(sweeps):
            ch.sweep(seed=4, sweep=2 + sw, precision=prec, dense=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        kms, nk = eng.kernel_time("hmm")
        eng.set_option("time_kernels", 0)
        _, g = ch.state_steps_last_sweep()
        esz = 8 if prec == "f64" else 4
        out[prec] = {"ms_per_sweep": 1e3 * dt / sweeps, "kernel_ms": kms / max(nk, 1), "mean_n": g / C,
                     "state_steps_per_sec": K * g * sweeps / dt, "failed": int((ch.status != 0).sum().item()),
                     "kernel_algorithmic_GBps": 2 * esz * K * g / (kms / max(nk, 1) / 1e3) / 1e9}
    eng.close()
    return {"config": "cfg4 dense K=%d exponential sMJP, %d chains, T=%g, Rao-Teh sweeps (candidates + dense FFBS)" % (K, C, T), **out}


def cfg4b(C=1024, K=64, T=1000.0, sweeps=1):
    """K=64 sMJP with GENERAL Weibull shapes over the long horizon of BASELINE configs[3] (the (v,l) path, run-time-K
    kernel ffbs_big.cuh with the chain state of the live window in a ring): fp32 with the exact window, fp64 with states
    below 2^-40 of the row sum dropped (the exact fp64 window does not fit on chip at K = 64)"""
    import torch
    from smjp_b200.chains import DeviceChains
    from smjp_b200.engine import Engine
    rng = np.random.default_rng(44)
    eng = Engine(0)
    eng.set_model(rng.uniform(0.6, 3.0, (K, K)), rng.uniform(0.7 * K, 1.3 * K, (K, K)), 2.0, emission(K), 1.0, T)
    obs_t = np.arange(0.5, T, 1.0)
    out = {}
    for prec, wl in (("f32", 0), ("f64", -40)):
        eng.set_option("window_log2", wl)
        ch = DeviceChains(eng, C, obs_t).init_prior(np.ones(K) / K, seed=1, honour_scale=True, cap_per_chain=int(4 * T) + 64)
        ch.simulate_observations(seed=2)
        ch.init_prior(np.ones(K) / K, seed=3, honour_scale=True, cap_per_chain=int(4 * T) + 64)
        ch.sweep(seed=4, sweep=0, precision=prec, mode="exact")
        torch.cuda.synchronize()
        eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs")
        t0 = time.perf_counter()
        for sw in range(sweeps):
            ch.sweep(seed=4, sweep=1 + sw, precision=prec, mode="exact")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        live = float(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs"))
        s, g = ch.state_steps_last_sweep()
        out[prec] = {"window": "exact" if wl == 0 else "pruned 2^%d" % wl, "ms_per_sweep": 1e3 * dt / sweeps, "mean_n": g / C,
                     "nominal_state_steps_per_sec": s * sweeps / dt, "retained_state_steps_per_sec": K * live / dt,
                     "live_fraction": K * live / max(s * sweeps, 1.0), "failed": int((ch.status != 0).sum().item())}
        del ch
    eng.set_option("window_log2", 0)
    eng.close()
    return {"config": "cfg4b K=%d general-shape sMJP, T=%g, %d chains (ffbs_big kernel, state ring over the live window)" % (K, T, C), **out}


def cfg5(C=16384, T=20.0, sweeps=4, threads=0):
    """10-state sMJP, many chains (one GPU's shard of the 65536-chain configuration)"""
    import torch
    from smjp_b200.chains import DeviceChains
    from smjp_b200.engine import Engine
    rng = np.random.default_rng(5)
    K = 10
    eng = Engine(0)
    eng.set_model(rng.uniform(0.6, 3.0, (K, K)), np.ones((K, K)), 2.0, emission(K), 1.0, T)
    eng.set_option("ffbs_threads", threads)
    obs_t = np.arange(0.1, T, 0.1)
    out = {"threads": threads}
    for prec in ("f64", "f32"):
        ch = DeviceChains(eng, C, obs_t).init_prior(np.ones(K) / K, seed=1).simulate_observations(seed=2)
        ch.init_prior(np.ones(K) / K, seed=3)
        for sw in range(3):
            ch.sweep(seed=4, sweep=sw, precision=prec)
        torch.cuda.synchronize()
        ss = 0.0
        t0 = time.perf_counter()
        for sw in range(sweeps):
            ch.sweep(seed=4, sweep=3 + sw, precision=prec)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        for _ in range(1):
            s, g = ch.state_steps_last_sweep()
        out[prec] = {"chain_sweeps_per_sec": C * sweeps / dt, "state_steps_per_sec": s * sweeps / dt, "mean_n": g / C,
                     "ms_per_sweep": 1e3 * dt / sweeps, "failed": int((ch.status != 0).sum().item()),
                     "algorithmic_GBps": (2 * (8 if prec == "f64" else 4) * s * sweeps / dt) / 1e9}
    eng.close()
    return {"config": "cfg5 shard: %d chains, K=10, T=%g" % (C, T), **out}


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg3", "cfg4", "cfg5"]
    for name in which:
        if name.startswith("cfg5:"):
            print(json.dumps(cfg5(threads=int(name[5:]))), flush=True)
        elif name.startswith("cfg3:"):  # cfg3:C[:cluster[:threads]] = filter runs in flight, launch geometry
            f = [int(x) for x in name[5:].split(":")] + [0, 0]
            print(json.dumps(cfg3(C=f[0], cluster=f[1], threads=f[2])), flush=True)
        else:
            print(json.dumps(globals()[name]()), flush=True)
