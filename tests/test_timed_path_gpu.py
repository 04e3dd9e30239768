"""GPU parity tests of the code path bench.py TIMES: scratch mode (no exported messages), so that the live window
opens and the backward pass reads g-rows (mode 1), on BASELINE-scale chains -- cfg2 (K=3, T=100, D=999, n ~ 540)
for ffbs_kernel<double,3,128,*>, ffbs_warp_kernel<float|double,3> and the fp32 block kernel, cfg5 (K=10, T=20,
n ~ 200) for ffbs_item_kernel -- against oracle.ffbs on the same grids and Philox uniforms (pytest -m gpu).

Reference lines: forward filter fast_hmm.py:202-335, backward sampler fast_hmm.py:104-191, compaction
smjp_utils.py:319-330, metrics mcmc_utils.py:40-56."""
import numpy as np
import pytest

from oracle import philox
from oracle import smjp_oracle as O

pytestmark = pytest.mark.gpu

SHAPE3 = np.array([[1.606, 1.933, 0.865], [1.938, 0.869, 1.751], [1.69, 0.64, 0.696]])  # experiments.py:172-174
SEED_BACK = 4


@pytest.fixture(scope="module")
def engine():
    from smjp_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


def make_chains(C, K, shape, t_end, seed0):
    """cfg2 / cfg5 style inputs per chain: observations simulated from a prior path (10:1 emission, every 0.1),
    grid = candidates of an INDEPENDENT prior path -- what a sweep of bench.py's chains sees"""
    scale = np.ones((K, K))
    emis = O.emission_matrix(K)
    obs_t = np.arange(0.1, t_end, 0.1)
    Ws, Ys = [], []
    for c in range(C):
        v, w = O.prior_path(shape, scale, np.ones(K) / K, t_end, seed=seed0 + 1, chain=c)
        Ys.append(O.simulate_observations(v, w, obs_t, emis, seed=seed0 + 2, chain=c))
        v, w = O.prior_path(shape, scale, np.ones(K) / K, t_end, seed=seed0 + 3, chain=c)
        Ws.append(O.candidates(v, w, shape, scale, 2.0, seed=seed0 + 4, sweep=0, chain=c))
    return scale, emis, obs_t, Ws, Ys


def oracle_runs(Ws, Ys, obs_t, shape, scale, emis, t_end, sweep):
    out = []
    for c, W in enumerate(Ws):
        n = len(W) - 1
        u = philox.u01(SEED_BACK, philox.TAG_BACKWARD, np.arange(n), 0, sweep, c)
        r = O.ffbs(W, obs_t, Ys[c], shape, scale, 2.0, emis, 1.0, t_end, u)
        r["u"] = u
        out.append(r)
    return out


def exact_live_pairs(alpha):
    """grid pairs (i, j) from the first grid point with a non-zero state to i, summed over the rows: what an exact
    live window that is re-examined at every step would process; also the widest window"""
    n, K, _ = alpha.shape
    nz = (alpha > 0).any(axis=1)
    first = np.array([int(np.argmax(nz[i])) for i in range(n)])
    width = np.arange(n) - first + 1
    return int(width.sum()), int(width.max())


def live_pairs(engine):
    return int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_live_pairs"))


def assert_exact_parity(r, oras, K):
    for c, ora in enumerate(oras):
        assert r["status"][c] == 0
        assert np.array_equal(r["aug_idx"][c], ora["samples"]), c
        assert np.array_equal(r["V"][c], ora["V"]) and np.array_equal(r["T"][c], ora["T"]), c
        tis, jm = O.chain_metrics(ora["V"], ora["T"], K)
        assert np.array_equal(r["jumps"][c], jm)
        assert np.abs(r["time_in_state"][c] - tis).max() < 1e-12
        assert np.abs(r["logZ"][c] - ora["logZ"]).max() < 1e-10, np.abs(r["logZ"][c] - ora["logZ"]).max()


def boundary_distance(ora, W, shape, scale, K, i):
    """distance of the uniform of draw i from the nearest boundary of the oracle's CDF at that draw, given the
    oracle's own sample at i + 1 (fast_hmm.py:153-181)"""
    n = len(W) - 1
    alpha, samples, u = ora["alpha"], ora["samples"], ora["u"]
    if i == n - 1:
        p = alpha[n - 1].reshape(-1)
    else:
        v_next, j_next = divmod(int(samples[i + 1]), n)
        w = np.zeros((K, n))
        if j_next == i + 1:
            tau = W[i + 1] - W[: i + 1]
            A = O.weibull_hazard(tau, shape, scale)
            w[:, : i + 1] = alpha[i, :, : i + 1] * (A[:, v_next, :] / (2.0 * A.sum(axis=1)))
        else:
            w[v_next, j_next] = 1.0
        p = w.reshape(-1)
    cdf = np.cumsum(p / p.sum())
    return float(np.min(np.abs(cdf - u[i])))


@pytest.fixture(scope="module")
def cfg2_case():
    C, K, t_end = 6, 3, 100.0
    scale, emis, obs_t, Ws, Ys = make_chains(C, K, SHAPE3, t_end, 100)
    oras = oracle_runs(Ws, Ys, obs_t, SHAPE3, scale, emis, t_end, sweep=7)
    return dict(C=C, K=K, t_end=t_end, scale=scale, emis=emis, obs_t=obs_t, Ws=Ws, Ys=Ys, oras=oras)


KERNEL_NAMES = {0: "ffbs_kernel", 1: "ffbs_item_kernel", 2: "ffbs_big_kernel", 3: "ffbs_warp_kernel", 4: "ffbs_fast_kernel",
                5: "ffbs_fastk_kernel"}


def set_kernel(engine, threads=0, warp=-1, fast=-1):
    engine.set_option("ffbs_threads", threads)
    engine.set_option("ffbs_warp", warp)
    engine.set_option("ffbs_fast", fast)


def kernel_used(engine):
    return KERNEL_NAMES[int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_kernel"))]


@pytest.mark.parametrize("threads,warp,fast,kernel", [
    (0, -1, -1, "ffbs_fast_kernel"),       # what a sweep launches by default: ffbs_fast_kernel<double,3,32,2,2,5> (warp per chain)
    (128, 0, 1, "ffbs_fast_kernel"), (64, 0, 1, "ffbs_fast_kernel"), (256, 0, 1, "ffbs_fast_kernel"), (32, 0, 1, "ffbs_fast_kernel"),
    (128, 0, 0, "ffbs_kernel"), (64, 0, 0, "ffbs_kernel"), (256, 0, 0, "ffbs_kernel"),   # the round-1 kernel (now: export / rescue)
    (0, 1, 0, "ffbs_warp_kernel")])
def test_cfg2_scratch_mode_fp64_equals_oracle(engine, cfg2_case, threads, warp, fast, kernel):
    """the instantiation the headline times (default options -> ffbs_fast_kernel<double,3,128,5>, scratch rows, open
    live window in a state ring), its other CTA sizes, the round-1 block kernel and the fp64 warp-per-chain ring
    kernel: bit-equal paths, logZ to 1e-10"""
    g = cfg2_case
    engine.set_model(SHAPE3, g["scale"], 2.0, g["emis"], 1.0, g["t_end"])
    set_kernel(engine, threads, warp, fast)
    try:
        live_pairs(engine)
        r = engine.ffbs_host(g["Ws"], [g["obs_t"]] * g["C"], g["Ys"], seed=SEED_BACK, sweep=7, precision="f64")
        live = live_pairs(engine)
        assert kernel_used(engine) == kernel
        # threads per chain: the fast kernel's default is one warp per chain (several chains per CTA)
        assert engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_threads") == (threads if threads and warp != 1 else 32 if kernel == "ffbs_fast_kernel" else 128)
        rescued = int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_rescued")) if kernel != "ffbs_kernel" else 0
    finally:
        set_kernel(engine)
    assert rescued == 0   # no chain was handed to the rescue pass: the kernel under test produced these results
    assert_exact_parity(r, g["oras"], g["K"])
    nominal = sum(n * (n + 1) // 2 for n in r["n"])
    exact = sum(exact_live_pairs(o["alpha"])[0] for o in g["oras"])
    widest = max(exact_live_pairs(o["alpha"])[1] for o in g["oras"])
    assert min(r["n"]) > 400 and widest < min(r["n"]) // 2      # BASELINE scale, and the window does open
    # the kernel drops only exact zeros and looks every few steps: never fewer pairs than the oracle's zero pattern
    # allows (deferred normalisation moves an underflow by a step or two at most: 1 % slack), and well below the
    # nominal triangle
    assert 0.99 * exact <= live < 0.75 * nominal, (exact, live, nominal)
    assert live < 1.1 * exact, (exact, live)


@pytest.mark.parametrize("split", [0, 1])
def test_cfg2_fast_kernel_fused_and_split_launches_equal_oracle(engine, cfg2_case, split):
    """the K <= 3 sweep kernel as ONE launch (forward + backward per chain) and as TWO (forward of every chain, then
    backward of every chain, scratch region per chain): the stat says which ran, both give the oracle's paths"""
    g = cfg2_case
    engine.set_model(SHAPE3, g["scale"], 2.0, g["emis"], 1.0, g["t_end"])
    engine.set_option("fast_split", split)
    try:
        r = engine.ffbs_host(g["Ws"], [g["obs_t"]] * g["C"], g["Ys"], seed=SEED_BACK, sweep=7, precision="f64")
        assert kernel_used(engine) == "ffbs_fast_kernel"
        assert int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"fast_split")) == split
        assert int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_rescued")) == 0
    finally:
        engine.set_option("fast_split", -1)
    assert_exact_parity(r, g["oras"], g["K"])


def test_fast_kernel_ring_overflow_and_range_go_to_the_rescue_pass(engine, cfg2_case):
    """results never depend on the ring size: with one ring slot (128 grid points, the cfg2 window needs ~250) every
    chain overflows and is redone by the rescue pass (ffbs.cuh) -- same paths; a grid with a 1e-300 gap leaves the
    range of the unclamped fp64 arithmetic and takes the same route"""
    g = cfg2_case
    engine.set_model(SHAPE3, g["scale"], 2.0, g["emis"], 1.0, g["t_end"])
    engine.set_option("fast_ring", 1)
    try:
        r = engine.ffbs_host(g["Ws"], [g["obs_t"]] * g["C"], g["Ys"], seed=SEED_BACK, sweep=7, precision="f64")
        assert kernel_used(engine) == "ffbs_fast_kernel"
        assert int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_rescued")) == g["C"]
    finally:
        engine.set_option("fast_ring", 0)
    assert_exact_parity(r, g["oras"], g["K"])
    W = np.sort(np.concatenate([g["Ws"][0], [g["Ws"][0][5] * (1 + 1e-15)]]))                 # a gap of ~1e-15: in range
    Wt = np.concatenate([[0.0, 1.0e-300], g["Ws"][1][1:]])                                   # and one of 1e-300
    us = [np.random.default_rng(1).uniform(size=len(x) - 1) for x in (W, Wt)]
    r = engine.ffbs_host([W, Wt], [g["obs_t"]] * 2, g["Ys"][:2], uniforms=us, precision="f64")
    assert int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_rescued")) == 1
    for c, (Wc, Y) in enumerate(zip((W, Wt), g["Ys"][:2])):
        ora = O.ffbs(Wc, g["obs_t"], Y, SHAPE3, g["scale"], 2.0, g["emis"], 1.0, g["t_end"], us[c])
        assert r["status"][c] == 0
        assert np.array_equal(r["V"][c], ora["V"]) and np.array_equal(r["T"][c], ora["T"])
        assert np.abs(r["logZ"][c] - ora["logZ"]).max() < 1e-9


@pytest.mark.parametrize("warp,fast", [(1, 0), (0, 0), (0, 1)])
def test_cfg2_scratch_mode_fp32_against_oracle(engine, cfg2_case, warp, fast):
    """ffbs_warp_kernel<float,3> (warp = 1: what bench.py's f32 block times) and the fp32 block kernel in scratch
    mode: logZ within 1e-4; a path may differ from the fp64 oracle's only where the uniform of the first differing
    draw sits within fp32 rounding of a CDF boundary -- counted and bounded"""
    g = cfg2_case
    K = g["K"]
    engine.set_model(SHAPE3, g["scale"], 2.0, g["emis"], 1.0, g["t_end"])
    set_kernel(engine, 0, warp, fast)
    try:
        live_pairs(engine)
        r = engine.ffbs_host(g["Ws"], [g["obs_t"]] * g["C"], g["Ys"], seed=SEED_BACK, sweep=7, precision="f32")
        live = live_pairs(engine)
        rescued = int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_rescued"))
        assert kernel_used(engine) == ("ffbs_warp_kernel" if warp else "ffbs_fast_kernel" if fast else "ffbs_kernel")
    finally:
        set_kernel(engine)
    assert np.all(r["status"] == 0)
    assert rescued == 0   # every chain was done by the fp32 kernel itself, not by the fp64 rescue pass
    nominal = sum(n * (n + 1) // 2 for n in r["n"])
    assert live < 0.45 * nominal, (live, nominal)   # fp32 underflows much earlier: the window is ~100 grid points
    differ = 0
    for c, ora in enumerate(g["oras"]):
        assert np.abs(r["logZ"][c] - ora["logZ"]).max() < 1e-4
        if np.array_equal(r["aug_idx"][c], ora["samples"]):
            assert np.array_equal(r["V"][c], ora["V"]) and np.array_equal(r["T"][c], ora["T"])
            continue
        differ += 1
        i = int(np.nonzero(r["aug_idx"][c] != ora["samples"])[0][-1])   # draws run last to first
        assert boundary_distance(ora, g["Ws"][c], SHAPE3, g["scale"], K, i) < 1e-4, (c, i)
    assert differ <= 1, differ


@pytest.mark.parametrize("fast,ways,kernel", [(-1, 0, "ffbs_fastk_kernel"), (1, 1, "ffbs_fastk_kernel"), (0, 0, "ffbs_item_kernel")])
def test_cfg5_scale_k10_scratch_mode_equals_oracle(engine, fast, ways, kernel):
    """K = 10, T = 20 (cfg5's chain shape, n ~ 200) in scratch mode with an open live window: ffbs_fastk_kernel<double,10>
    (what a sweep launches by default; one and two slots in flight) and ffbs_item_kernel<double,10,...> (export / rescue
    kernel): bit-equal paths, logZ to 1e-10; fp32 within tolerance"""
    C, K, t_end = 6, 10, 20.0
    shape = np.random.default_rng(5).uniform(0.6, 3.0, (K, K))
    scale, emis, obs_t, Ws, Ys = make_chains(C, K, shape, t_end, 500)
    oras = oracle_runs(Ws, Ys, obs_t, shape, scale, emis, t_end, sweep=2)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    set_kernel(engine, 0, -1, fast)
    engine.set_option("fast_ways", ways)
    try:
        live_pairs(engine)
        r = engine.ffbs_host(Ws, [obs_t] * C, Ys, seed=SEED_BACK, sweep=2, precision="f64")
        live = live_pairs(engine)
        assert kernel_used(engine) == kernel
        rescued = int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_rescued")) if fast else 0
        r32 = engine.ffbs_host(Ws, [obs_t] * C, Ys, seed=SEED_BACK, sweep=2, precision="f32")
        assert kernel_used(engine) == kernel
        rescued32 = int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_rescued"))
    finally:
        set_kernel(engine)
        engine.set_option("fast_ways", 0)
    assert rescued == 0 and rescued32 == 0
    assert_exact_parity(r, oras, K)
    nominal = sum(n * (n + 1) // 2 for n in r["n"])
    exact = sum(exact_live_pairs(o["alpha"])[0] for o in oras)
    assert min(r["n"]) > 120
    assert 0.99 * exact <= live < 0.85 * nominal, (exact, live, nominal)
    assert np.all(r32["status"] == 0)
    differ = 0
    for c, ora in enumerate(oras):
        assert np.abs(r32["logZ"][c] - ora["logZ"]).max() < 2e-4
        if not np.array_equal(r32["aug_idx"][c], ora["samples"]):
            differ += 1
            i = int(np.nonzero(r32["aug_idx"][c] != ora["samples"])[0][-1])
            assert boundary_distance(ora, Ws[c], shape, scale, K, i) < 1e-4, (c, i)
    assert differ <= 1, differ


@pytest.mark.parametrize("K", [4, 5, 6, 7, 8, 9, 10])
def test_fastk_every_K_against_oracle(engine, K):
    """ffbs_fastk_kernel for every K it is instantiated for (JPS = 32 / K grid points per warp slot: 8, 6, 5, 4, 4, 3,
    3), ragged batch, non-unit scales, Omega = 3, a ring too small for one of the chains (rescue pass)"""
    rng = np.random.default_rng(70 + K)
    t_end, omega = 12.0, 3.0
    shape = rng.uniform(0.6, 3.0, (K, K)); scale = rng.uniform(2.0, 4.0, (K, K)); emis = O.emission_matrix(K)
    engine.set_model(shape, scale, omega, emis, 0.8, t_end)
    Ws = [np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, m)), [t_end]]) for m in (150, 6, 3, 77, 230)]
    obs_t = np.arange(0.25, t_end, 0.25); obs_y = rng.integers(0, K, len(obs_t))
    us = [rng.uniform(size=len(W) - 1) for W in Ws]
    r = engine.ffbs_host(Ws, obs_t, obs_y, uniforms=us, precision="f64")
    assert kernel_used(engine) == "ffbs_fastk_kernel"
    for c, W in enumerate(Ws):
        ora = O.ffbs(W, obs_t, obs_y, shape, scale, omega, emis, 0.8, t_end, us[c])
        assert r["status"][c] == 0
        assert np.array_equal(r["aug_idx"][c], ora["samples"]), c
        assert np.array_equal(r["V"][c], ora["V"]) and np.array_equal(r["T"][c], ora["T"])
        assert np.abs(r["logZ"][c] - ora["logZ"]).max() < 1e-10
        tis, jm = O.chain_metrics(ora["V"], ora["T"], K)
        assert np.array_equal(r["jumps"][c], jm) and np.abs(r["time_in_state"][c] - tis).max() < 1e-12


def test_cfg2_sweep_device_loop_equals_oracle_composition(engine):
    """three whole sweeps (candidates -> FFBS, device Philox) of cfg2-scale chains through DeviceChains.sweep -- the
    loop bench.py times -- against the oracle's composition, stage by stage"""
    from smjp_b200.chains import DeviceChains
    C, K, t_end = 4, 3, 100.0
    scale, emis = np.ones((K, K)), O.emission_matrix(K)
    obs_t = np.arange(0.1, t_end, 0.1)
    engine.set_model(SHAPE3, scale, 2.0, emis, 1.0, t_end)
    ch = DeviceChains(engine, C, obs_t).init_prior(np.ones(K) / K, seed=1).simulate_observations(seed=2)
    Y = ch.obs_y.cpu().numpy().reshape(C, -1)
    ch.init_prior(np.ones(K) / K, seed=3)
    V, T = ch.paths_host()
    for sw in range(3):
        ch.sweep(seed=9, sweep=sw)
        Vd, Td = ch.paths_host()
        Wd = ch.grids_host()
        assert int((ch.status != 0).sum().item()) == 0
        tis, jm = ch.time_in_state.cpu().numpy(), ch.jumps.cpu().numpy()
        for c in range(C):
            W = O.candidates(V[c], T[c], SHAPE3, scale, 2.0, seed=9, sweep=sw, chain=c)
            assert len(W) == len(Wd[c]) and np.allclose(W, Wd[c], rtol=1e-12, atol=1e-14)
            u = philox.u01(9, philox.TAG_BACKWARD, np.arange(len(W) - 1), 0, sw, c)
            ora = O.ffbs(Wd[c], obs_t, Y[c], SHAPE3, scale, 2.0, emis, 1.0, t_end, u)  # the device's grid, bit for bit
            assert np.array_equal(Vd[c], ora["V"]) and np.array_equal(Td[c], ora["T"]), (sw, c)
            t_o, j_o = O.chain_metrics(ora["V"], ora["T"], K)
            assert np.array_equal(jm[c], j_o) and np.abs(tis[c] - t_o).max() < 1e-12
            V[c], T[c] = Vd[c], Td[c]


@pytest.mark.parametrize("K,precision,warp,fast", [(3, "f64", 0, 1), (3, "f64", 0, 0), (3, "f64", 1, 0), (2, "f64", 0, 1), (2, "f64", 0, 0),
                                                   (5, "f64", 0, 0), (5, "f64", 0, 1), (10, "f64", 0, 0), (10, "f64", 0, 1), (12, "f64", 0, 0),
                                                   (3, "f32", 1, 0), (3, "f32", 0, 0), (3, "f32", 0, 1), (10, "f32", 0, 0), (10, "f32", 0, 1)])
def test_grid_points_within_isclose_of_t_end(engine, K, precision, warp, fast):
    """A candidate within numpy.isclose() of t_end (|w - T| <= 1e-8 + 1e-5 T; 0.5 % of the cfg2 chain-sweeps) gives two
    or more rows without the B_v factor in the emission (smjp_utils.py:544-547) while the transition keeps
    A_vv'/B_v (:409-412): every kernel, scratch and export mode, against the oracle (itself pinned on the reference's
    own output for this case: tests/golden/ffbs_final_rows_*.npz)"""
    from smjp_b200.engine import triangular_to_dense
    rng = np.random.default_rng(40 + K)
    t_end, n0 = 30.0, 150
    shape = rng.uniform(0.6, 3.0, (K, K)); scale = np.ones((K, K)); emis = O.emission_matrix(K)
    engine.set_model(shape, scale, 2.0, emis, 1.0, t_end)
    inner = np.sort(rng.uniform(0, t_end - 0.01, n0))
    Ws = [np.concatenate([[0.0], inner, [t_end - 2.0e-4, t_end]]),                      # two final rows
          np.concatenate([[0.0], inner[::2], [t_end - 2.5e-4, t_end - 1.0e-4, t_end - 3.0e-5, t_end]]),   # four
          np.concatenate([[0.0], inner[::3], [t_end]])]                                   # the usual single one
    obs_t = np.arange(0.1, t_end, 0.1); obs_y = rng.integers(0, K, len(obs_t))
    us = [rng.uniform(size=len(W) - 1) for W in Ws]
    set_kernel(engine, 0, warp, fast)
    try:
        q = engine.ffbs_host(Ws, obs_t, obs_y, uniforms=us, precision=precision)                       # scratch mode
        r = engine.ffbs_host(Ws, obs_t, obs_y, uniforms=us, precision=precision, want_messages=True)   # export mode
    finally:
        set_kernel(engine)
    tol = 1e-10 if precision == "f64" else 2e-4
    for c, W in enumerate(Ws):
        m = len(W) - 1
        ora = O.ffbs(W, obs_t, obs_y, shape, scale, 2.0, emis, 1.0, t_end, us[c])
        for res in (q, r):
            assert res["status"][c] == 0
            assert np.abs(res["logZ"][c] - ora["logZ"]).max() < tol, (c, np.abs(res["logZ"][c] - ora["logZ"]).max())
            if precision == "f64":
                assert np.array_equal(res["aug_idx"][c], ora["samples"])
                assert np.array_equal(res["V"][c], ora["V"]) and np.array_equal(res["T"][c], ora["T"])
        dense = triangular_to_dense(r["messages"][c], m, K).astype(np.float64)
        ref = ora["alpha"].reshape(m, K * m)
        big = ref > (1e-200 if precision == "f64" else 1e-6)
        assert (np.abs(dense[big] - ref[big]) / ref[big]).max() < (1e-9 if precision == "f64" else 2e-4)
