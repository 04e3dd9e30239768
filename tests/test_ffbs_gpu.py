"""GPU parity tests of the FFBS kernel, through the C ABI, against the reference golden fixtures
and the oracle (run on the B200 box: pytest -m gpu)."""
import numpy as np
import pytest

from conftest import ffbs_fixture_names, load_golden
from oracle import smjp_oracle as O

pytestmark = pytest.mark.gpu

REL_F64 = 1e-6   # north_star tolerance for fp64 messages / log-likelihoods
REL_F32 = 1e-4   # north_star tolerance for fp32


@pytest.fixture(scope="module")
def engine():
    from smjp_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


def run_fixture(engine, g, precision, **kw):
    from smjp_b200.engine import triangular_to_dense
    K = g["shape"].shape[0]
    engine.set_model(g["shape"], g["scale"], float(g["omega"]), O.emission_matrix(K), float(g["power"]),
                     float(g["t_end"]))
    r = engine.ffbs_host([g["W"]], [g["obs_t"]], [g["obs_y"]], uniforms=[g["uniforms"]], precision=precision,
                         want_messages=True, **kw)
    n = len(g["W"]) - 1
    r["dense"] = triangular_to_dense(r["messages"][0], n, K)
    return r


def rel_err(a, ref, floor):
    """max relative error over entries that carry mass (> floor); absolute error elsewhere"""
    big = ref > floor
    rel = np.abs(a[big] - ref[big]) / ref[big]
    ab = np.abs(a[~big] - ref[~big])
    return (rel.max() if rel.size else 0.0), (ab.max() if ab.size else 0.0)


@pytest.mark.parametrize("name", ffbs_fixture_names())
def test_fp64_matches_reference_fixture(engine, name):
    g = load_golden("ffbs_" + name)
    r = run_fixture(engine, g, "f64")
    assert r["status"][0] == 0
    rel, ab = rel_err(r["dense"], g["alpha"], 1e-300)
    assert rel < REL_F64, rel
    assert rel < 1e-11  # what fp64 actually achieves
    assert np.array_equal(r["dense"] > 0, g["alpha"] > 0)
    # bit-exact backward path given identical host uniforms
    assert np.array_equal(r["aug_idx"][0], g["samples"])
    assert np.array_equal(r["V"][0], g["V"])
    assert np.array_equal(r["T"][0], g["T"])
    # log-normalisers: unpinned by the reference, compared with the oracle
    K = g["shape"].shape[0]
    ora = O.ffbs(g["W"], g["obs_t"], g["obs_y"], g["shape"], g["scale"], float(g["omega"]),
                 O.emission_matrix(K), float(g["power"]), float(g["t_end"]), g["uniforms"])
    assert np.allclose(r["logZ"][0], ora["logZ"], rtol=REL_F64, atol=1e-12)
    t, j = O.chain_metrics(g["V"], g["T"], K)
    assert np.array_equal(r["jumps"][0], j)
    assert np.allclose(r["time_in_state"][0], t, rtol=0, atol=1e-14)


@pytest.mark.parametrize("name", ffbs_fixture_names())
def test_fp32_matches_reference_fixture(engine, name):
    g = load_golden("ffbs_" + name)
    r = run_fixture(engine, g, "f32")
    assert r["status"][0] == 0
    rel, ab = rel_err(r["dense"].astype(np.float64), g["alpha"], 1e-6)
    assert rel < REL_F32, rel
    assert ab < 1e-9
    K = g["shape"].shape[0]
    ora = O.ffbs(g["W"], g["obs_t"], g["obs_y"], g["shape"], g["scale"], float(g["omega"]),
                 O.emission_matrix(K), float(g["power"]), float(g["t_end"]), g["uniforms"])
    assert np.allclose(r["logZ"][0], ora["logZ"], rtol=REL_F32, atol=1e-5)


def random_problem(rng, K, t_end, n, D, omega=2.0):
    shape = rng.uniform(0.6, 3.0, (K, K))
    scale = np.ones((K, K))
    W = np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, n - 1)), [t_end]])
    obs_t = np.sort(rng.uniform(0, t_end, D))
    obs_y = rng.integers(0, K, D)
    return shape, scale, W, obs_t, obs_y


@pytest.mark.parametrize("K,n,D", [(2, 40, 10), (3, 200, 150), (5, 130, 64), (10, 90, 30), (3, 700, 999)])
def test_fp64_matches_oracle_random(engine, K, n, D):
    from smjp_b200.engine import triangular_to_dense
    rng = np.random.default_rng(100 * K + n)
    t_end = n / 5.0
    shape, scale, W, obs_t, obs_y = random_problem(rng, K, t_end, n, D)
    emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    C = 3
    us = [rng.uniform(size=n) for _ in range(C)]
    r = engine.ffbs_host([W] * C, obs_t, obs_y, uniforms=us, precision="f64", want_messages=True)
    assert np.all(r["status"] == 0)
    for c in range(C):
        ora = O.ffbs(W, obs_t, obs_y, shape, scale, 2.0, emis, 1.0, t_end, us[c])
        dense = triangular_to_dense(r["messages"][c], n, K)
        rel, ab = rel_err(dense, ora["alpha"].reshape(n, K * n), 1e-250)
        assert rel < 1e-9, rel
        assert np.allclose(r["logZ"][c], ora["logZ"], rtol=1e-10, atol=1e-12)
        assert np.array_equal(r["aug_idx"][c], ora["samples"])
        assert np.array_equal(r["V"][c], ora["V"])
        assert np.array_equal(r["T"][c], ora["T"])


@pytest.mark.parametrize("K,n,D,prec", [(11, 70, 30, "f64"), (16, 150, 60, "f64"), (33, 60, 25, "f64"),
                                        (64, 40, 20, "f64"), (24, 90, 40, "f32"), (64, 100, 50, "f32")])
def test_large_state_space_kernel_matches_oracle(engine, K, n, D, prec):
    """K > 10 with general shapes runs on the run-time-K kernel (ffbs_big.cuh): exported messages, paths
    with host uniforms, and the scratch (g-row) path with device Philox"""
    from oracle import philox
    from smjp_b200.engine import triangular_to_dense
    rng = np.random.default_rng(1000 * K + n)
    # fp32 has no room for a whole row below 1e-38 of the previous one (long intervals at large K): finer grid
    t_end = n / (5.0 if prec == "f64" else 40.0)
    shape, scale, W, obs_t, obs_y = random_problem(rng, K, t_end, n, D)
    emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    W2 = np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, n // 2 - 1)), [t_end]])  # ragged batch
    us = [rng.uniform(size=n), rng.uniform(size=len(W2) - 1)]
    r = engine.ffbs_host([W, W2], obs_t, obs_y, uniforms=us, precision=prec, want_messages=True)
    q = engine.ffbs_host([W, W2], obs_t, obs_y, uniforms=us, precision=prec, want_messages=False)
    p = engine.ffbs_host([W, W2], obs_t, obs_y, seed=99, sweep=3, precision=prec)
    assert np.all(r["status"] == 0) and np.all(q["status"] == 0) and np.all(p["status"] == 0)
    for c, Wc in enumerate([W, W2]):
        m = len(Wc) - 1
        ora = O.ffbs(Wc, obs_t, obs_y, shape, scale, 2.0, emis, 1.0, t_end, us[c])
        dense = triangular_to_dense(r["messages"][c], m, K)
        # normalisation is deferred by one step, so an entry below ~1e-308 / Z_i of the row mass flushes to
        # zero one step before it does in the reference (K = 16: Z_i reaches 1e-120): floor 1e-180
        rel, ab = rel_err(dense, ora["alpha"].reshape(m, K * m), 1e-180 if prec == "f64" else 1e-30)
        assert rel < (1e-9 if prec == "f64" else 2e-4), rel
        assert np.allclose(r["logZ"][c], ora["logZ"], rtol=1e-10 if prec == "f64" else 1e-4, atol=1e-12 if prec == "f64" else 1e-4)
        if prec == "f64":
            for res in (r, q):  # exported (alpha rows) and scratch (g rows) backward passes agree with the oracle
                assert np.array_equal(res["aug_idx"][c], ora["samples"])
                assert np.array_equal(res["V"][c], ora["V"]) and np.array_equal(res["T"][c], ora["T"])
            tis, jm = O.chain_metrics(ora["V"], ora["T"], K)
            assert np.abs(q["time_in_state"][c] - tis).max() < 1e-12 and np.array_equal(q["jumps"][c], jm)
            u = philox.u01(99, philox.TAG_BACKWARD, np.arange(m), 0, 3, c)
            orap = O.ffbs(Wc, obs_t, obs_y, shape, scale, 2.0, emis, 1.0, t_end, u)
            assert np.array_equal(p["V"][c], orap["V"]) and np.array_equal(p["T"][c], orap["T"])


def test_large_state_space_sweeps_and_capacity_error(engine):
    """Rao-Teh sweeps at K = 12 through DeviceChains; exported messages of a grid whose state does not fit on chip are refused"""
    from smjp_b200._lib import EngineError
    from smjp_b200.chains import DeviceChains
    rng = np.random.default_rng(5)
    K, t_end = 12, 4.0
    shape = rng.uniform(0.6, 3.0, (K, K)); scale = rng.uniform(2.0, 6.0, (K, K))
    engine.set_model(shape, scale, 2.0, O.emission_matrix(K), 1.0, t_end)
    ch = DeviceChains(engine, 64, np.arange(0.25, t_end, 0.25)).init_prior(np.ones(K) / K, seed=1, honour_scale=True)
    ch.simulate_observations(seed=2)
    for sw in range(3):
        ch.sweep(seed=3, sweep=sw)
    assert int((ch.status != 0).sum().item()) == 0
    V, T = ch.paths_host()
    assert all(t[0] == 0.0 and t[-1] == t_end and np.all(np.diff(t) > 0) for t in T)
    K = 64
    engine.set_model(np.full((K, K), 1.5), np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, 1.0)
    W = np.linspace(0.0, 1.0, 2001)
    # scratch mode: the chain state lives in a ring over the live window, so the length of the grid is no limit any
    # more; a window that outgrows the ring would end the chain with SMJP_ST_OVERFLOW, never with wrong numbers
    r = engine.ffbs_host([W], np.zeros(0), np.zeros(0, np.int32), uniforms=[np.full(2000, 0.5)])
    assert r["status"][0] in (0, 4)
    with pytest.raises((EngineError, RuntimeError, ValueError)):  # exported messages need the whole triangle on chip
        engine.ffbs_host([W], np.zeros(0), np.zeros(0, np.int32), uniforms=[np.full(2000, 0.5)], want_messages=True)


def test_ragged_batch_and_philox_backward(engine):
    """chains of different lengths in one launch; device Philox uniforms reproduce the oracle's."""
    from oracle import philox
    rng = np.random.default_rng(7)
    K, t_end = 3, 6.0
    shape = rng.uniform(0.6, 3.0, (K, K)); scale = np.ones((K, K)); emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    Ws, ots, oys = [], [], []
    for c, n in enumerate([1, 2, 5, 33, 64, 65, 129, 17]):
        Ws.append(np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, n - 1)), [t_end]]))
        D = int(rng.integers(0, 30))
        ots.append(np.sort(rng.uniform(0, t_end, D))); oys.append(rng.integers(0, K, D))
    r = engine.ffbs_host(Ws, ots, oys, uniforms=None, seed=1234, sweep=5, precision="f64")
    assert np.all(r["status"] == 0)
    for c, W in enumerate(Ws):
        n = len(W) - 1
        u = philox.u01(1234, philox.TAG_BACKWARD, np.arange(n), 0, 5, c)
        ora = O.ffbs(W, ots[c], oys[c], shape, scale, 2.0, emis, 1.0, t_end, u)
        assert np.array_equal(r["aug_idx"][c], ora["samples"])
        assert np.array_equal(r["V"][c], ora["V"]) and np.array_equal(r["T"][c], ora["T"])
        assert np.allclose(r["logZ"][c], ora["logZ"], rtol=1e-10, atol=1e-12)


def test_non_increasing_grid_is_reported(engine):
    from smjp_b200.engine import raise_for_status
    K = 3
    engine.set_model(np.full((K, K), 1.2), np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, 2.0)
    r = engine.ffbs_host([np.array([0.0, 0.5, 0.5, 2.0]), np.array([0.0, 1.0, 2.0])], np.zeros(0), np.zeros(0, int),
                         uniforms=[np.full(3, 0.5), np.full(2, 0.5)])
    assert r["status"][0] == 1 and r["status"][1] == 0
    with pytest.raises(ValueError):
        raise_for_status(r["status"])


def test_parameter_range_of_the_device_exponent_arithmetic(engine):
    """shapes whose hazard exponents would leave the integer range of the fp64 2^x are refused up front;
    wide but valid shapes (k = 0.05 .. 40) over a grid with 1e-9 gaps still match the oracle"""
    from smjp_b200._lib import EngineError
    K = 3
    for bad in (1.0e4, 1.0e-5):
        shape = np.full((K, K), 1.5); shape[1, 2] = bad
        with pytest.raises((EngineError, ValueError)):
            engine.set_model(shape, np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, 2.0)
    rng = np.random.default_rng(3)
    shape = np.array([[0.05, 1.0, 40.0], [3.0, 0.3, 12.0], [25.0, 0.7, 1.9]])
    scale = rng.uniform(0.8, 2.5, (K, K)); emis = O.emission_matrix(K); t_end = 3.0
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    inner = np.sort(rng.uniform(0.0, t_end, 40))
    W = np.concatenate([[0.0], inner, inner[::5] + 1e-9, [t_end]]); W.sort()
    ot = np.sort(rng.uniform(0, t_end, 12)); oy = rng.integers(0, K, 12)
    u = rng.uniform(size=len(W) - 1)
    r = engine.ffbs_host([W], ot, oy, uniforms=[u], precision="f64", want_messages=True)
    ora = O.ffbs(W, ot, oy, shape, scale, 2.0, emis, 1.0, t_end, u)
    assert r["status"][0] == 0
    assert np.allclose(r["logZ"][0], ora["logZ"], rtol=1e-9, atol=1e-11)
    assert np.array_equal(r["V"][0], ora["V"]) and np.array_equal(r["T"][0], ora["T"])


def test_fp32_path_mismatches_only_near_cdf_boundaries(engine):
    """fp32 messages differ by ~1e-5, so a sampled index may differ from the fp64 one only when the
    uniform lands within that tolerance of a CDF boundary."""
    rng = np.random.default_rng(11)
    K, n, t_end = 3, 120, 24.0
    shape, scale, W, obs_t, obs_y = random_problem(rng, K, t_end, n, 80)
    engine.set_model(shape, scale, 2.0, O.emission_matrix(K), 1.0, t_end)
    C = 64
    us = [rng.uniform(size=n) for _ in range(C)]
    r64 = engine.ffbs_host([W] * C, obs_t, obs_y, uniforms=us, precision="f64")
    r32 = engine.ffbs_host([W] * C, obs_t, obs_y, uniforms=us, precision="f32")
    same = sum(np.array_equal(a, b) for a, b in zip(r64["aug_idx"], r32["aug_idx"]))
    assert same >= C - 2, same


def test_product_observation_likelihood_matches_oracle(engine):
    """obs_combine='product' (not the reference's sum) against the oracle's same switch"""
    from smjp_b200.engine import triangular_to_dense
    rng = np.random.default_rng(21)
    K, n, t_end = 3, 30, 6.0
    shape, scale, W, obs_t, obs_y = random_problem(rng, K, t_end, n, 120)
    emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 0.7, t_end, obs_combine="product")
    u = rng.uniform(size=n)
    r = engine.ffbs_host([W], [obs_t], [obs_y], uniforms=[u], want_messages=True)
    ora = O.ffbs(W, obs_t, obs_y, shape, scale, 2.0, emis, 0.7, t_end, u, obs_combine="product")
    rel, ab = rel_err(triangular_to_dense(r["messages"][0], n, K), ora["alpha"].reshape(n, K * n), 1e-250)
    assert rel < 1e-9 and np.array_equal(r["aug_idx"][0], ora["samples"])
    engine.set_model(shape, scale, 2.0, emis, 0.7, t_end)  # back to the reference behaviour
    r2 = engine.ffbs_host([W], [obs_t], [obs_y], uniforms=[u], want_messages=True)
    assert np.abs(r2["messages"][0] - r["messages"][0]).max() > 1e-3  # the two targets really differ


def test_fp32_chain_that_underflows_is_redone_in_fp64(engine):
    """a grid interval over which every state's survival factor is below fp32's range (exp(-200)) leaves the
    fp32 pass without a row sum; by default such chains are redone by the fp64 kernel in the same call"""
    rng = np.random.default_rng(21)
    K, t_end = 3, 4.0
    shape = np.full((K, K), 2.5); scale = np.full((K, K), 0.25)
    emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    fine = np.sort(rng.uniform(1.0, t_end, 60))
    W_bad = np.concatenate([[0.0], fine, [t_end]])          # first interval [0, ~1]: H_v grows by ~100
    W_ok = np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, 400)), [t_end]])  # ~0.01 apart: fine in fp32
    obs_t = np.arange(0.5, t_end, 0.5); obs_y = rng.integers(0, K, len(obs_t))
    us = [rng.uniform(size=len(W_bad) - 1), rng.uniform(size=len(W_ok) - 1)]
    ref64 = engine.ffbs_host([W_bad, W_ok], obs_t, obs_y, uniforms=us, precision="f64")
    assert np.all(ref64["status"] == 0)
    try:
        engine.set_option("f32_rescue", 0)
        raw = engine.ffbs_host([W_bad, W_ok], obs_t, obs_y, uniforms=us, precision="f32")
        assert raw["status"][0] == 2 and raw["status"][1] == 0  # SMJP_ST_DEGENERATE
    finally:
        engine.set_option("f32_rescue", 1)
    r = engine.ffbs_host([W_bad, W_ok], obs_t, obs_y, uniforms=us, precision="f32")
    assert np.all(r["status"] == 0)
    assert np.array_equal(r["V"][0], ref64["V"][0]) and np.array_equal(r["T"][0], ref64["T"][0])  # the fp64 result
    assert np.array_equal(r["time_in_state"][0], ref64["time_in_state"][0])
    assert len(r["V"][1]) >= 2 and r["T"][1][0] == 0.0


def test_empty_batch_and_observations_on_grid_points(engine):
    """C = 0 is a no-op through the C ABI; an observation whose time equals a grid point belongs to BOTH
    neighbouring intervals, as in the reference (closed interval test, smjp_utils.py:45)."""
    import ctypes
    K, t_end = 3, 4.0
    rng = np.random.default_rng(11)
    shape = rng.uniform(0.6, 3.0, (K, K)); scale = np.ones((K, K)); emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    rc = engine.lib.smjp_ffbs_host(engine.ctx, 0, None, None, None, None, None, None, ctypes.c_uint64(0), 0, 0, None, None,
                                   None, None, None, None, None, None, None, None)
    assert rc == 0
    W = np.array([0.0, 0.5, 1.0, 1.75, 2.5, 3.0, t_end])
    ot = np.array([0.5, 1.0, 1.2, 2.5, 2.5, 4.0])      # four on grid points (one of them twice), one at t_end
    oy = np.array([0, 1, 2, 1, 0, 2])
    u = rng.uniform(size=len(W) - 1)
    r = engine.ffbs_host([W], ot, oy, uniforms=[u], precision="f64", want_messages=True)
    ora = O.ffbs(W, ot, oy, shape, scale, 2.0, emis, 1.0, t_end, u)
    assert r["status"][0] == 0
    assert np.allclose(r["logZ"][0], ora["logZ"], rtol=1e-10, atol=1e-12)
    assert np.array_equal(r["aug_idx"][0], ora["samples"])
    assert np.array_equal(r["V"][0], ora["V"]) and np.array_equal(r["T"][0], ora["T"])
    # the shared observation really is counted twice: dropping it from either side changes logZ of both intervals
    r2 = engine.ffbs_host([W], ot[1:], oy[1:], uniforms=[u], precision="f64")
    assert abs(r2["logZ"][0][0] - r["logZ"][0][0]) > 1e-6 and abs(r2["logZ"][0][1] - r["logZ"][0][1]) > 1e-6


@pytest.mark.parametrize("K,precision", [(3, "f64"), (3, "f32"), (6, "f64"), (12, "f64")])
def test_observation_symbol_out_of_range_is_flagged_on_the_device(engine, K, precision):
    """the device entry points cannot afford a host pass over the observations: every kernel checks the symbol before it
    indexes the emission table and flags the chain (SMJP_ST_BAD_INPUT -> ValueError); the other chains are unaffected"""
    import torch
    from smjp_b200.engine import raise_for_status
    rng = np.random.default_rng(K)
    t_end, n = 5.0, 40
    shape, scale, W, obs_t, obs_y = random_problem(rng, K, t_end, n, 30)
    engine.set_model(shape, scale, 2.0, O.emission_matrix(K), 1.0, t_end)
    dev = torch.device("cuda", engine.device)
    C = 3
    d = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x, dtype=dt)).to(dev)
    ys = np.tile(obs_y, C).astype(np.int32)
    ys[len(obs_y) + 7] = K + 5      # chain 1 carries a bad symbol
    tot = C * (n + 1)
    out = dict(pv=torch.zeros(tot, dtype=torch.int32, device=dev), pt=torch.zeros(tot, dtype=torch.float64, device=dev),
               pl=torch.zeros(C, dtype=torch.int32, device=dev), st=torch.zeros(C, dtype=torch.int32, device=dev))
    for scratch in (True, False):
        msgs = None if scratch else torch.zeros(C * K * n * (n + 1) // 2, dtype=torch.float64 if precision == "f64" else torch.float32, device=dev)
        moffs = None if scratch else d(np.arange(C + 1) * (K * n * (n + 1) // 2), np.int64)
        engine.ffbs_dev(C, d(np.arange(C + 1) * (n + 1), np.int64), d(np.tile(W, C), np.float64), d(np.arange(C + 1) * len(obs_t), np.int64),
                        d(np.tile(obs_t, C), np.float64), d(ys, np.int32), out["pv"], out["pt"], out["pl"], out["st"], n, tot,
                        seed=3, precision=precision, messages=msgs, msg_offs=moffs)
        engine.sync()
        st = out["st"].cpu().numpy()
        assert st[0] == 0 and st[2] == 0 and (st[1] & 8) and out["pl"].cpu().numpy()[1] == 0
        with pytest.raises(ValueError):
            raise_for_status(st)


def _long_grid_case(K, n, seed):
    rng = np.random.default_rng(seed)
    t_end = n / 4.0
    # total leaving rate ~ 4 per unit time, grid spacing 1/4: a state loses ~2^-4 per grid step (thinning 1/2 x survival),
    # so the exact fp64 window is ~270 grid points and the 2^-70 one ~20
    shape = rng.uniform(0.6, 3.0, (K, K)); scale = rng.uniform(0.15 * K, 0.35 * K, (K, K))
    W = np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, n - 1)), [t_end]])
    obs_t = np.arange(0.25, t_end, 0.25); obs_y = rng.integers(0, K, len(obs_t))
    return shape, scale, W, obs_t, obs_y, t_end, rng.uniform(size=n)


def test_large_state_space_long_grid_exact_window(engine):
    """ffbs_big.cuh keeps the chain state of the LIVE window in a ring (it used to hold the whole grid: K = 24 stopped
    near n = 500): K = 24, n = 700, exact fp64 window, scratch mode -- oracle parity, and the window really opened"""
    K, n = 24, 700
    shape, scale, W, obs_t, obs_y, t_end, u = _long_grid_case(K, n, 1)
    emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_live_pairs")
    r = engine.ffbs_host([W], obs_t, obs_y, uniforms=[u], precision="f64")
    live = engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_live_pairs")
    ora = O.ffbs(W, obs_t, obs_y, shape, scale, 2.0, emis, 1.0, t_end, u)
    assert r["status"][0] == 0
    assert np.array_equal(r["aug_idx"][0], ora["samples"])
    assert np.array_equal(r["V"][0], ora["V"]) and np.array_equal(r["T"][0], ora["T"])
    assert np.abs(r["logZ"][0] - ora["logZ"]).max() < 1e-10
    assert live < 0.8 * n * (n + 1) / 2
    with pytest.raises(Exception):  # exported messages need the whole triangle on chip: refused, as before
        engine.ffbs_host([W], obs_t, obs_y, uniforms=[u], precision="f64", want_messages=True)


def test_k64_long_grid_pruned_window_against_the_exact_oracle(engine):
    """BASELINE configs[3]'s general-shape sMJP at K = 64 on a LONG grid (n = 520; the first version stopped at 140):
    the exact fp64 window (hundreds of grid points: a state loses about the thinning factor 1/2 per grid step relative
    to the row) does not fit the ring at K = 64 (two slots of 64 grid points), the pruned one (states below 2^-40 of the
    row sum dropped, SURVEY 7 hard part 2) does: same path as the exact oracle, log Z within 1e-9; without pruning the
    chain is refused (SMJP_ST_OVERFLOW), never silently truncated; fp32's exact window fits its five slots"""
    from smjp_b200._lib import ST_OVERFLOW
    K, n = 64, 520
    shape, scale, W, obs_t, obs_y, t_end, u = _long_grid_case(K, n, 2)
    emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    ora = O.ffbs(W, obs_t, obs_y, shape, scale, 2.0, emis, 1.0, t_end, u)
    raw = engine.ffbs_host([W], obs_t, obs_y, uniforms=[u], precision="f64")
    assert raw["status"][0] & ST_OVERFLOW
    engine.set_option("window_log2", -40)
    try:
        r = engine.ffbs_host([W], obs_t, obs_y, uniforms=[u], precision="f64")
    finally:
        engine.set_option("window_log2", 0)
    assert r["status"][0] == 0
    assert np.array_equal(r["V"][0], ora["V"]) and np.array_equal(r["T"][0], ora["T"])
    assert np.abs(r["logZ"][0] - ora["logZ"]).max() < 1e-9
    r32 = engine.ffbs_host([W], obs_t, obs_y, uniforms=[u], precision="f32")
    assert r32["status"][0] == 0 and np.abs(r32["logZ"][0] - ora["logZ"]).max() < 5e-4
