"""How many entries of the fp32 / fp64 forward messages of a cfg2-like chain are exactly zero (dead states)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import smjp_oracle as O
from smjp_b200.engine import Engine
from smjp_b200.chains import DeviceChains
eng = Engine(0)
K, T = 3, 100.0
shape = np.array([[1.606, 1.933, 0.865], [1.938, 0.869, 1.751], [1.69, 0.64, 0.696]])
eng.set_model(shape, np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, T)
obs_t = np.arange(0.1, T, 0.1)
ch = DeviceChains(eng, 8, obs_t).init_prior(np.ones(K) / K, seed=1).simulate_observations(seed=2)
ch.init_prior(np.ones(K) / K, seed=3)
for sw in range(5):
    ch.sweep(seed=4, sweep=sw)
W = ch.grids_host()
oy = ch.obs_y.cpu().numpy().reshape(8, -1)
for prec in ("f32", "f64"):
    for c in range(2):
        r = eng.ffbs_host([W[c]], obs_t, oy[c], seed=1, precision=prec, want_messages=True)
        m = r["messages"][0]
        n = len(W[c]) - 1
        # per row statistics: entries below thresholds relative to the row max
        z = float(np.mean(m == 0))
        tiny = {t: float(np.mean(m < t)) for t in (1e-30, 1e-20, 1e-12)}
        print(prec, "chain", c, "n", n, "exact zeros", round(z, 4), "below", tiny, flush=True)

# live-window accounting of the sweep kernel (scratch mode)
ch2 = DeviceChains(eng, 512, obs_t).init_prior(np.ones(K) / K, seed=1).simulate_observations(seed=2)
ch2.init_prior(np.ones(K) / K, seed=3)
for prec in ("f64", "f32"):
    for sw in range(3):
        ch2.sweep(seed=4, sweep=sw, precision=prec)
    eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs")
    ch2.sweep(seed=4, sweep=9, precision=prec)
    live = eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs")
    s, g = ch2.state_steps_last_sweep()
    print(prec, "live pairs", live, "nominal pairs", s / K, "live fraction", live / (s / K), flush=True)
