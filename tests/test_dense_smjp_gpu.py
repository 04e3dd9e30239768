"""GPU parity tests of the K-state collapse of the shape == 1 sMJP (smjp_ffbs_dense_dev,
smjp_sweep_dense_dev) against the oracle twin and against the augmented (v,l) FFBS
(pytest -m gpu)."""
import numpy as np
import pytest

from oracle import philox
from oracle import smjp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from smjp_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


def make_case(rng, K, n, t_end, D):
    scale = rng.uniform(0.5, 2.0, (K, K))
    W = np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, n - 1)), [t_end]])
    obs_t = np.sort(rng.uniform(0, t_end, D))
    obs_y = rng.integers(0, K, D)
    return scale, W, obs_t, obs_y


@pytest.mark.parametrize("K,n,prec", [(3, 40, "f64"), (5, 120, "f64"), (33, 60, "f64"), (64, 150, "f64"),
                                      (10, 90, "f32"), (64, 80, "f32")])
def test_dense_ffbs_matches_oracle_with_host_uniforms(engine, K, n, prec):
    rng = np.random.default_rng(K * 100 + n)
    omega, t_end = 2.0, 9.0
    scale, _, _, _ = make_case(rng, K, n, t_end, 10)
    emis = O.emission_matrix(K)
    engine.set_model(np.ones((K, K)), scale, omega, emis, 1.0, t_end)
    C = 6
    Ws, ots, oys, us, uj = [], [], [], [], []
    for c in range(C):
        _, W, obs_t, obs_y = make_case(rng, K, n + 7 * c, t_end, 10 + 5 * c)
        Ws.append(W); ots.append(obs_t); oys.append(obs_y)
        us.append(rng.uniform(size=len(W) - 1)); uj.append(rng.uniform(size=len(W) - 1))
    r = engine.ffbs_dense(Ws, ots, oys, u_state=us, u_jump=uj, precision=prec)
    assert np.all(r["status"] == 0)
    for c in range(C):
        o = O.dense_ffbs_collapsed(Ws[c], ots[c], oys[c], scale, omega, emis, 1.0, t_end, us[c], uj[c])
        tol = 1e-10 if prec == "f64" else 2e-4  # fp32 tolerance on the per-step log normaliser
        assert np.abs(r["logZ"][c] - o["logZ"]).max() < tol
        if prec == "f64":  # the path is index work: bit-exact
            assert np.array_equal(r["V"][c], o["V"])
            assert np.array_equal(r["T"][c], o["T"])
            tis, jm = O.chain_metrics(o["V"], o["T"], K)
            assert np.abs(r["time_in_state"][c] - tis).max() < 1e-12
            assert np.array_equal(r["jumps"][c], jm)


def test_dense_ffbs_philox_streams_match_oracle(engine):
    rng = np.random.default_rng(11)
    K, n, omega, t_end, seed, sweep = 7, 80, 2.5, 6.0, 0xABCDEF12345, 9
    scale, _, _, _ = make_case(rng, K, n, t_end, 10)
    emis = O.emission_matrix(K)
    engine.set_model(np.ones((K, K)), scale, omega, emis, 1.0, t_end)
    engine.set_option("chain_base", 5)
    try:
        cases = [make_case(rng, K, n + c, t_end, 30)[1:] for c in range(4)]
        r = engine.ffbs_dense([c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases], seed=seed, sweep=sweep)
    finally:
        engine.set_option("chain_base", 0)
    for c, (W, obs_t, obs_y) in enumerate(cases):
        m = len(W) - 1
        us = [float(philox.u01(seed, philox.TAG_BACKWARD, 2 * i, 0, sweep, 5 + c)) for i in range(m)]
        uj = [float(philox.u01(seed, philox.TAG_BACKWARD, 2 * i + 1, 0, sweep, 5 + c)) for i in range(m)]
        o = O.dense_ffbs_collapsed(W, obs_t, obs_y, scale, omega, emis, 1.0, t_end, us, uj)
        assert np.array_equal(r["V"][c], o["V"]) and np.array_equal(r["T"][c], o["T"])


def test_dense_ffbs_product_observations_and_final_step(engine):
    """obs_combine='product', several observations per interval, a grid whose last interval ends at t_end"""
    rng = np.random.default_rng(12)
    K, omega, t_end = 4, 2.0, 3.0
    scale = rng.uniform(0.5, 2.0, (K, K)); emis = O.emission_matrix(K)
    W = np.array([0.0, 0.4, 1.7, 2.2, t_end])
    obs_t = np.array([0.1, 0.2, 0.4, 1.0, 2.9, 3.0]); obs_y = np.array([0, 1, 1, 2, 3, 3])
    us, uj = rng.uniform(size=4), rng.uniform(size=4)
    for comb in ("sum", "product"):
        engine.set_model(np.ones((K, K)), scale, omega, emis, 0.5, t_end, obs_combine=comb)
        r = engine.ffbs_dense([W], obs_t, obs_y, u_state=[us], u_jump=[uj])
        o = O.dense_ffbs_collapsed(W, obs_t, obs_y, scale, omega, emis, 0.5, t_end, us, uj, obs_combine=comb)
        assert np.abs(r["logZ"][0] - o["logZ"]).max() < 1e-12
        assert np.array_equal(r["V"][0], o["V"]) and np.array_equal(r["T"][0], o["T"])
    engine.set_model(np.ones((K, K)), scale, omega, emis, 1.0, t_end)


def test_dense_ffbs_rejects_non_exponential_models_and_bad_grids(engine):
    from smjp_b200._lib import EngineError
    K = 3
    engine.set_model(np.full((K, K), 2.0), np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, 2.0)
    with pytest.raises((ValueError, EngineError)):
        engine.ffbs_dense([np.array([0.0, 1.0, 2.0])], np.zeros(0), np.zeros(0, np.int32))
    engine.set_model(np.ones((K, K)), np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, 2.0)
    r = engine.ffbs_dense([np.array([0.0, 1.0, 1.0, 2.0]), np.array([0.0, 1.0, 2.0])], np.zeros(0), np.zeros(0, np.int32))
    assert r["status"][0] & 1 and r["status"][1] == 0  # smjp_utils.py:398-399 as a per-chain status bit


def test_dense_and_augmented_ffbs_sample_the_same_posterior(engine):
    """shape == 1: the (v,l) FFBS and its K-state collapse draw paths from one distribution
    (different uniform streams, so the comparison is distributional: per-state time and entry counts)."""
    from scipy import stats
    rng = np.random.default_rng(13)
    K, n, omega, t_end, C = 3, 30, 2.0, 5.0, 6000
    scale, W, obs_t, obs_y = make_case(rng, K, n, t_end, 12)
    emis = O.emission_matrix(K)
    engine.set_model(np.ones((K, K)), scale, omega, emis, 1.0, t_end)
    a = engine.ffbs_host([W] * C, obs_t, obs_y, seed=21, want_logZ=False, want_aug_idx=False)
    d = engine.ffbs_dense([W] * C, obs_t, obs_y, seed=22)
    assert np.all(a["status"] == 0) and np.all(d["status"] == 0)
    for s in range(K):
        assert stats.ks_2samp(a["time_in_state"][:, s], d["time_in_state"][:, s]).pvalue > 1e-3
        assert stats.ks_2samp(a["jumps"][:, s], d["jumps"][:, s]).pvalue > 1e-3
    la = np.array([len(v) for v in a["V"]]); ld = np.array([len(v) for v in d["V"]])
    assert stats.ks_2samp(la, ld).pvalue > 1e-3


def test_dense_sweeps_reach_the_same_posterior_as_augmented_sweeps(engine):
    from scipy import stats
    from smjp_b200.chains import DeviceChains
    rng = np.random.default_rng(14)
    K, omega, t_end, C = 4, 2.0, 4.0, 4000
    scale = rng.uniform(0.5, 2.0, (K, K)); emis = O.emission_matrix(K)
    engine.set_model(np.ones((K, K)), scale, omega, emis, 1.0, t_end)
    obs_t = np.arange(0.25, t_end, 0.5)
    y = rng.integers(0, K, len(obs_t)).astype(np.int32)
    out = {}
    for dense in (False, True):
        ch = DeviceChains(engine, C, obs_t).init_prior(np.ones(K) / K, seed=31 + dense)
        ch.set_observations([y] * C)
        for sw in range(12):
            ch.sweep(seed=41 + dense, sweep=sw, dense=dense)
        assert int((ch.status != 0).sum().item()) == 0
        out[dense] = (ch.time_in_state.cpu().numpy(), ch.jumps.cpu().numpy())
    for s in range(K):
        assert stats.ks_2samp(out[False][0][:, s], out[True][0][:, s]).pvalue > 1e-3
        assert stats.ks_2samp(out[False][1][:, s], out[True][1][:, s]).pvalue > 1e-3


def test_dense_sweep_k64_runs_and_is_seed_reproducible(engine):
    from smjp_b200.chains import DeviceChains
    rng = np.random.default_rng(15)
    K, omega, t_end, C = 64, 2.0, 6.0, 64
    scale = rng.uniform(20.0, 80.0, (K, K)); emis = O.emission_matrix(K)
    engine.set_model(np.ones((K, K)), scale, omega, emis, 1.0, t_end)
    obs_t = np.arange(0.1, t_end, 0.1)
    runs = []
    for rep in range(2):
        ch = DeviceChains(engine, C, obs_t).init_prior(np.ones(K) / K, seed=1).simulate_observations(seed=2)
        for sw in range(3):
            ch.sweep(seed=3, sweep=sw, dense=True, precision="f32" if rep == 2 else "f64")
        assert int((ch.status != 0).sum().item()) == 0
        runs.append(ch.paths_host())
    for c in range(C):
        assert np.array_equal(runs[0][0][c], runs[1][0][c]) and np.array_equal(runs[0][1][c], runs[1][1][c])
        assert runs[0][1][c][0] == 0.0 and runs[0][1][c][-1] == t_end and np.all(np.diff(runs[0][1][c]) > 0)


@pytest.mark.parametrize("K,C", [(64, 300), (40, 150), (64, 1)])
def test_tensor_core_forward_matches_oracle_and_the_warp_kernel(engine, K, C):
    """dense_tc.cuh: the forward matvecs of 128 chains at a time as [128 x 64] x [64 x 64] tcgen05 products (3xTF32,
    A operand in tensor memory).  Ragged tile (C is no multiple of 128), chains of different lengths in one tile, padded
    states (K = 40): per-step log normalisers against the fp64 oracle within the fp32 tolerance, and the same sampled
    paths as the warp-per-chain kernel for (nearly) every chain (same uniforms, fp32 messages differing at the 1e-6
    level)."""
    rng = np.random.default_rng(K + C)
    omega, t_end = 2.0, 9.0
    scale = rng.uniform(0.5 * K / 8, 2.0 * K / 8, (K, K))   # leaving rates of order 8 per unit time
    emis = O.emission_matrix(K)
    engine.set_model(np.ones((K, K)), scale, omega, emis, 1.0, t_end)
    Ws, ots, oys, us, uj = [], [], [], [], []
    for c in range(C):
        n = int(rng.integers(20, 90))
        W = np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, n - 1)), [t_end]])
        D = int(rng.integers(0, 25))
        Ws.append(W); ots.append(np.sort(rng.uniform(0, t_end, D))); oys.append(rng.integers(0, K, D))
        us.append(rng.uniform(size=n)); uj.append(rng.uniform(size=n))
    out = {}
    for tc in (1, 0):
        engine.set_option("dense_tc", tc)
        try:
            out[tc] = engine.ffbs_dense(Ws, ots, oys, u_state=us, u_jump=uj, precision="f32")
            assert engine.lib.smjp_ctx_get_stat(engine.ctx, b"dense_tc") == tc
        finally:
            engine.set_option("dense_tc", -1)
    r, w = out[1], out[0]
    assert np.all(r["status"] == 0) and np.all(w["status"] == 0)
    same = 0
    for c in range(C):
        if c < 12:
            o = O.dense_ffbs_collapsed(Ws[c], ots[c], oys[c], scale, omega, emis, 1.0, t_end, us[c], uj[c])
            assert np.abs(r["logZ"][c] - o["logZ"]).max() < 2e-4
        assert np.abs(r["logZ"][c] - w["logZ"][c]).max() < 1e-4
        same += int(np.array_equal(r["V"][c], w["V"][c]) and np.array_equal(r["T"][c], w["T"][c]))
    assert same >= C - max(1, C // 50), (same, C)
