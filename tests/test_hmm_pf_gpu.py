"""GPU parity tests: dense K-state HMM (legacy hidden_markov_model API) and the bootstrap particle
filter, against the reference fixture and the oracle (pytest -m gpu)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import smjp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from smjp_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


def test_dense_hmm_matches_reference_fixture(engine):
    from smjp_b200.engine import hmm_dense_host
    g = load_golden("hmm_dense")
    r = hmm_dense_host(engine, g["P"], g["pi0"], emis=g["E"], y=[g["data"]], uniforms=[g["uniforms"]])
    assert r["status"][0] == 0
    log_alpha = np.log(r["alpha_hat"][0]) + np.cumsum(r["logc"][0])[:, None]  # the reference's unnormalised log messages
    assert np.abs(log_alpha - g["log_alphas"]).max() < 1e-10
    assert abs(np.exp(r["loglik"][0]) - float(g["prob"])) / float(g["prob"]) < 1e-10
    assert np.array_equal(r["states"][0], g["samples"])


@pytest.mark.parametrize("S,n,prec,tol", [(3, 50, "f64", 1e-10), (10, 300, "f64", 1e-10), (64, 200, "f64", 1e-10),
                                          (100, 40, "f64", 1e-10), (64, 200, "f32", 1e-4)])
def test_dense_hmm_matches_oracle(engine, S, n, prec, tol):
    from smjp_b200.engine import hmm_dense_host
    rng = np.random.default_rng(S * 1000 + n)
    P = rng.dirichlet(np.ones(S), size=S)
    pi0 = rng.dirichlet(np.ones(S))
    C = 5
    Es = [rng.uniform(0.05, 1.0, (n + c, S)) for c in range(C)]
    us = [rng.uniform(size=n + c) for c in range(C)]
    r = hmm_dense_host(engine, P, pi0, E=Es, uniforms=us, precision=prec)
    assert np.all(r["status"] == 0)
    for c in range(C):
        la, ah, logc, ll = O.hmm_forward_dense(P, Es[c], pi0)
        assert np.abs(r["alpha_hat"][c] - ah).max() < tol
        assert abs(r["loglik"][c] - ll) < max(tol * abs(ll), 1e-9)
        if prec == "f64":
            assert np.array_equal(r["states"][c], O.hmm_backward_dense(ah, P, us[c]))


def test_dense_hmm_time_varying_transition(engine):
    from smjp_b200.engine import hmm_dense_host
    rng = np.random.default_rng(3)
    S, n = 6, 30
    P = rng.dirichlet(np.ones(S), size=(n, S))
    pi0 = np.ones(S) / S
    E = rng.uniform(0.1, 1, (n, S)); u = rng.uniform(size=n)
    r = hmm_dense_host(engine, P, pi0, E=[E], uniforms=[u])
    la, ah, logc, ll = O.hmm_forward_dense(P, E, pi0)
    assert np.abs(r["alpha_hat"][0] - ah).max() < 1e-12
    assert np.array_equal(r["states"][0], O.hmm_backward_dense(ah, P, u))


def test_collapsed_smjp_equals_augmented_filter_on_gpu(engine):
    """shape == 1: marginalising the (v,l) FFBS messages over l equals the dense K-state filter (A.5)."""
    from smjp_b200.engine import hmm_dense_host, triangular_to_dense
    rng = np.random.default_rng(4)
    K, n, omega, t_end = 5, 60, 2.0, 12.0
    shape = np.ones((K, K)); scale = rng.uniform(0.5, 2.0, (K, K)); emis = O.emission_matrix(K)
    W = np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, n - 1)), [t_end]])
    obs_t = np.sort(rng.uniform(0, t_end, 40)); obs_y = rng.integers(0, K, 40)
    engine.set_model(shape, scale, omega, emis, 1.0, t_end)
    r = engine.ffbs_host([W], [obs_t], [obs_y], uniforms=[rng.uniform(size=n)], want_messages=True)
    marg = triangular_to_dense(r["messages"][0], n, K).reshape(n, K, n).sum(axis=2)
    A_v = (1.0 / scale).sum(axis=1)
    E = np.zeros((n, K))
    for i in range(n):
        lx = O.obs_factor(obs_t, obs_y, emis, 1.0, W[i], W[i + 1])
        E[i] = lx * (omega * A_v if i < n - 1 else 1.0) * np.exp(-omega * A_v * (W[i + 1] - W[i]))
    d = hmm_dense_host(engine, O.collapse_matrix(shape, scale, omega), np.ones(K) / K, E=[E], sample=False)
    assert np.abs(d["alpha_hat"][0] - marg).max() < 1e-12
    assert np.abs(d["logc"][0] - r["logZ"][0]).max() < 1e-10


@pytest.mark.parametrize("resampler,extend", [("systematic", False), ("multinomial", False), ("systematic", True)])
def test_particle_filter_matches_oracle(engine, resampler, extend):
    from smjp_b200.engine import pf_host
    K, t_end, N = 3, 2.0, 96
    shape = np.array([[1.606, 1.933, 0.865], [1.938, 0.869, 1.751], [1.69, 0.64, 0.696]])
    scale = np.array([[1.0, 0.8, 1.3], [0.7, 1.0, 1.1], [1.2, 0.9, 1.0]])
    emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    obs_t = np.arange(0.1, 2.0, 0.1)
    y = np.array([1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 3, 3, 2, 1, 1, 3]) - 1
    C = 6
    r = pf_host(engine, obs_t, y, np.ones(K) / K, N, seed=17, resampler=resampler, C=C, extend=extend)
    assert np.all(r["status"] == 0)
    for c in range(C):
        V, T, ll, Z = O.particle_filter(obs_t, y, shape, scale, emis, np.ones(K) / K, t_end, N, seed=17, chain=c,
                                        resampler=resampler, extend=extend)
        assert np.allclose(r["logZ"][c], Z, rtol=1e-12, atol=1e-12)
        assert abs(r["loglik"][c] - ll) < 1e-10
        assert np.array_equal(r["V"][c], V)
        assert np.allclose(r["T"][c], T, rtol=1e-12, atol=0)


def test_particle_filter_large_and_unbiased(engine):
    """65 536 particles (cfg3 scale): log-normaliser estimates from independent runs agree tightly and
    match a small-N ensemble average (unbiasedness of exp(sum Z_d))."""
    from smjp_b200.engine import pf_host
    rng = np.random.default_rng(2)
    K, t_end = 5, 10.0
    shape = rng.uniform(0.6, 3.0, (K, K)); scale = np.ones((K, K)); emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    obs_t = np.arange(0.1, t_end, 0.1); y = rng.integers(0, K, len(obs_t))
    big = pf_host(engine, obs_t, y, np.ones(K) / K, 65536, seed=5, C=4)
    assert np.all(big["status"] == 0)
    assert np.ptp(big["loglik"]) < 0.2
    small = pf_host(engine, obs_t, y, np.ones(K) / K, 512, seed=6, C=256)
    m = np.max(small["loglik"])
    avg = m + np.log(np.mean(np.exp(small["loglik"] - m)))
    assert abs(avg - np.mean(big["loglik"])) < 0.35
    for c in range(4):
        assert big["T"][c][0] == 0.0 and big["T"][c][-1] == t_end and np.all(np.diff(big["T"][c]) > 0)


def test_particle_filter_paths_follow_the_exact_posterior(engine):
    """shape == 1 (a Markov jump process): posterior time-in-state has an exact answer from a fine-grid
    forward-backward; the PF path sample means must match it for BOTH resamplers.  (Systematic
    offspring are ordered by ancestor: returning particle 0, as the reference does after multinomial
    resampling, would be biased -- the kernel draws the returned index uniformly in that mode.)"""
    from scipy.linalg import expm
    from smjp_b200.engine import pf_host
    K, t_end = 3, 2.0
    emis = O.emission_matrix(K)
    engine.set_model(np.ones((K, K)), np.ones((K, K)), 2.0, emis, 1.0, t_end)
    obs_t = np.arange(0.1, 2.0, 0.1)
    y = np.array([1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 3, 3, 2, 1, 1, 3]) - 1
    Q = np.ones((K, K)); np.fill_diagonal(Q, 0); Q -= np.diag(Q.sum(1))
    dt = 0.001; n = int(round(t_end / dt)); P = expm(Q * dt)
    oi = {int(round(t / dt)): v for t, v in zip(obs_t, y)}
    al = np.zeros((n + 1, K)); a = np.ones(K) / K
    for t in range(n + 1):
        a = a @ P if t else a
        a = a * emis[:, oi[t]] if t in oi else a
        a /= a.sum(); al[t] = a
    be = np.zeros((n + 1, K)); b = np.ones(K)
    for t in range(n, -1, -1):
        be[t] = b
        b = P @ (b * (emis[:, oi[t]] if t in oi else 1.0)); b /= b.sum()
    post = al * be; post /= post.sum(1, keepdims=True)
    exact = post[:-1].sum(0) * dt
    for res in ("systematic", "multinomial"):
        r = pf_host(engine, obs_t, y, np.ones(K) / K, 4096, seed=31, resampler=res, C=3000)
        tt = np.array([O.chain_metrics(V, T, K)[0] for V, T in zip(r["V"], r["T"])])
        assert np.abs(tt.mean(0) - exact).max() < 0.02, (res, tt.mean(0), exact)


def test_gamma_hazard_particle_filter_matches_oracle(engine):
    """BASELINE configs[2]: gamma holding times (new; the reference has Weibull hazards only).  Same Philox
    streams as the oracle, whose conditional gamma draw inverts scipy's regularised incomplete gamma."""
    from smjp_b200.engine import pf_host
    rng = np.random.default_rng(33)
    K, t_end, N = 5, 3.0, 64
    shape = rng.uniform(0.6, 3.0, (K, K)); scale = rng.uniform(0.5, 1.5, (K, K))
    emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    obs_t = np.arange(0.1, t_end, 0.1); y = rng.integers(0, K, len(obs_t))
    r = pf_host(engine, obs_t, y, np.ones(K) / K, N, seed=3, C=4, family="gamma")
    assert np.all(r["status"] == 0)
    for c in range(4):
        V, T, ll, Z = O.particle_filter(obs_t, y, shape, scale, emis, np.ones(K) / K, t_end, N, seed=3, chain=c,
                                        family="gamma")
        assert np.allclose(r["logZ"][c], Z, rtol=1e-10, atol=1e-10)
        assert np.array_equal(r["V"][c], V) and np.allclose(r["T"][c], T, rtol=1e-9, atol=0)
    # and the draws really are gamma: exponential special case a = 1 equals the Weibull k = 1 filter
    engine.set_model(np.ones((K, K)), scale, 2.0, emis, 1.0, t_end)
    g = pf_host(engine, obs_t, y, np.ones(K) / K, 256, seed=4, C=2, family="gamma")
    w = pf_host(engine, obs_t, y, np.ones(K) / K, 256, seed=4, C=2, family="weibull")
    assert np.allclose(g["loglik"], w["loglik"], rtol=1e-9) and all(np.array_equal(a, b) for a, b in zip(g["V"], w["V"]))
    with pytest.raises(Exception):
        engine.set_option("hazard_family", 1)
        try:
            engine.ffbs_host([np.array([0.0, 1.0, t_end])], np.zeros(0), np.zeros(0, int), uniforms=[np.array([0.5, 0.5])])
        finally:
            engine.set_option("hazard_family", 0)


@pytest.mark.parametrize("cs", [1, 2, 8, 16])
def test_particle_filter_cluster_sizes_agree(engine, cs):
    """the filter run split over a thread-block cluster (DSMEM totals + cluster barriers) is the same
    filter: identical ancestors and paths for every cluster size, equal to the oracle"""
    from smjp_b200.engine import pf_host
    K, t_end, N = 3, 2.0, 4096
    shape = np.array([[1.606, 1.933, 0.865], [1.938, 0.869, 1.751], [1.69, 0.64, 0.696]])
    emis = O.emission_matrix(K)
    engine.set_model(shape, np.ones((K, K)), 2.0, emis, 1.0, t_end)
    obs_t = np.arange(0.1, 2.0, 0.1)
    y = np.array([1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 3, 3, 2, 1, 1, 3]) - 1
    engine.set_option("pf_cluster", 1)
    ref = pf_host(engine, obs_t, y, np.ones(K) / K, N, seed=8, C=3)
    engine.set_option("pf_cluster", cs)
    try:
        r = pf_host(engine, obs_t, y, np.ones(K) / K, N, seed=8, C=3)
    finally:
        engine.set_option("pf_cluster", 0)
    assert np.all(r["status"] == 0)
    assert np.allclose(r["loglik"], ref["loglik"], rtol=0, atol=1e-10)
    for c in range(3):
        assert np.array_equal(r["V"][c], ref["V"][c]) and np.array_equal(r["T"][c], ref["T"][c])


@pytest.mark.parametrize("family", ["weibull", "gamma"])
def test_particle_filter_at_cfg3_scale_reproduces_the_exact_mjp_likelihood(engine, family):
    """Statistical anchor at BASELINE configs[2]'s scale: N = 65536 particles, K = 5, systematic resampling, cluster
    launch (the runs are spread over several CTAs each).  With exponential holding times (Weibull shape 1 == gamma
    shape 1) the sMJP is a Markov jump process and log p(y) is exact: pi0' prod_d [expm(Q dt_d) diag(P(y_d | .))] 1.
    The filter's log-normaliser (pmcmc.py:294-297, summed) must reproduce it -- for BOTH hazard families, through the
    host and the device-pointer entry points."""
    import torch
    from scipy.linalg import expm
    from smjp_b200.engine import pf_dev, pf_host
    rng = np.random.default_rng(65536)
    K, t_end, N, C = 5, 10.0, 65536, 6
    scale = rng.uniform(1.0, 4.0, (K, K))
    emis = O.emission_matrix(K)
    engine.set_model(np.ones((K, K)), scale, 2.0, emis, 1.0, t_end)
    obs_t = np.arange(0.4, t_end, 0.4)
    y = rng.integers(0, K, len(obs_t))
    pi0 = np.array([0.1, 0.3, 0.2, 0.25, 0.15])
    Q = 1.0 / scale
    np.fill_diagonal(Q, 0.0)            # self-jumps do not move the state
    Q -= np.diag(Q.sum(axis=1))
    a, ll, t_prev = pi0.copy(), 0.0, 0.0
    for t, s in zip(obs_t, y):
        a = (a @ expm(Q * (t - t_prev))) * emis[:, s]
        ll += np.log(a.sum())
        a /= a.sum()
        t_prev = t
    r = pf_host(engine, obs_t, y, pi0, N, seed=11, C=C, family=family)
    assert np.all(r["status"] == 0)
    assert engine.lib.smjp_ctx_get_stat(engine.ctx, b"pf_cluster") > 1        # the cluster path
    assert np.abs(r["loglik"] - ll).max() < 0.15, (r["loglik"], ll)            # sd of one estimate ~ 0.035 at this N
    assert abs(r["loglik"].mean() - ll) < 0.06                                  # sd of the mean of 6 ~ 0.015
    # device-pointer entry point: same runs, same numbers
    dev = torch.device("cuda", engine.device)
    D = len(obs_t)
    d = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x, dtype=dt)).to(dev)
    cap = 256
    out = dict(ll=torch.zeros(C, dtype=torch.float64, device=dev), pv=torch.zeros(C * cap, dtype=torch.int32, device=dev),
               pt=torch.zeros(C * cap, dtype=torch.float64, device=dev), pl=torch.zeros(C, dtype=torch.int32, device=dev),
               st=torch.zeros(C, dtype=torch.int32, device=dev))
    pf_dev(engine, C, d(np.arange(C + 1) * D, np.int64), d(np.tile(obs_t, C), np.float64), d(np.tile(y, C), np.int32),
           d(pi0, np.float64), N, 11, out["ll"], d(np.arange(C + 1) * cap, np.int64), out["pv"], out["pt"], out["pl"], out["st"], D,
           family=family)
    engine.sync()
    assert np.array_equal(out["ll"].cpu().numpy(), r["loglik"])
    assert int(out["st"].abs().sum().item()) == 0
