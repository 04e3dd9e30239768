"""Latency of the host-buffer entry points on a single tiny chain (cfg1 scale): where does a sweep's time go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import smjp_oracle as O
from smjp_b200.engine import Engine, candidates_host, sweep_host

eng = Engine(0)
K, T = 3, 2.0
shape = np.array([[1.606, 1.933, 0.865], [1.938, 0.869, 1.751], [1.69, 0.64, 0.696]])
eng.set_model(shape, np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, T)
obs_t = np.arange(0.1, T, 0.1); obs_y = np.zeros(len(obs_t), np.int32)
V, Tt = [np.array([0, 1, 1])], [np.array([0.0, 0.9, 2.0])]
W = candidates_host(eng, V, Tt, seed=1)
def timeit(f, n=200):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    return 1e3 * (time.perf_counter() - t0) / n
print("candidates_host  ms", round(timeit(lambda: candidates_host(eng, V, Tt, seed=1)), 3))
print("ffbs_host        ms", round(timeit(lambda: eng.ffbs_host(W, obs_t, obs_y, seed=1)), 3))
print("sweep_host       ms", round(timeit(lambda: sweep_host(eng, V, Tt, obs_t, obs_y, seed=1)), 3))
print("set_model        ms", round(timeit(lambda: eng.set_model(shape, np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, T)), 3))
