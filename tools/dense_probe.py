#!/usr/bin/env python
"""cfg4 dense K=64 path: warp-per-chain kernel vs tensor-core forward (dense_tc.cuh) at several chain counts.
    python tools/dense_probe.py C:T [C:T ...]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(C, T, K=64, sweeps=3):
    import torch
    from smjp_b200.chains import DeviceChains
    from smjp_b200.engine import Engine
    rng = np.random.default_rng(4)
    E = np.ones((K, K)) + 9 * np.eye(K)
    E /= E.sum(axis=1, keepdims=True)
    eng = Engine(0)
    eng.set_model(np.ones((K, K)), rng.uniform(0.5 * K, 1.5 * K, (K, K)), 2.0, E, 1.0, T)
    obs_t = np.arange(0.5, T, 1.0)
    out = {"C": C, "T": T, "K": K}
    for tc in (0, 1):
        eng.set_option("dense_tc", tc)
        ch = DeviceChains(eng, C, obs_t).init_prior(np.ones(K) / K, seed=1, honour_scale=True, cap_per_chain=max(64, int(8 * T))).simulate_observations(seed=2)
        ch.init_prior(np.ones(K) / K, seed=3, honour_scale=True, cap_per_chain=max(64, int(8 * T)))
        for sw in range(2):
            ch.sweep(seed=4, sweep=sw, precision="f32", dense=True)
        torch.cuda.synchronize()
        eng.set_option("time_kernels", 1)
        eng.kernel_time("hmm")
        t0 = time.perf_counter()
        for sw in range(sweeps):
            ch.sweep(seed=4, sweep=2 + sw, precision="f32", dense=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        kms, nk = eng.kernel_time("hmm")
        eng.set_option("time_kernels", 0)
        _, g = ch.state_steps_last_sweep()
        out["tc" if tc else "warp"] = {"ms_per_sweep": 1e3 * dt / sweeps, "ffbs_ms": kms / max(nk, 1), "mean_n": g / C,
                                       "failed": int((ch.status != 0).sum().item()),
                                       "tflops_ffbs": 2.0 * K * K * g / (kms / max(nk, 1) / 1e3) / 1e12,
                                       "used_tc": int(eng.lib.smjp_ctx_get_stat(eng.ctx, b"dense_tc"))}
        del ch
        torch.cuda.empty_cache()
    eng.close()
    return out


if __name__ == "__main__":
    for a in sys.argv[1:] or ["1024:2000"]:
        C, T = a.split(":")
        print(json.dumps(run(int(C), float(T))), flush=True)
