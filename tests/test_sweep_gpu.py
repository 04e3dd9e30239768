"""GPU parity tests: candidates, prior sampler, observation simulator and the full sweep against the
oracle's Philox-exact restatement (pytest -m gpu on the B200 box)."""
import numpy as np
import pytest

from oracle import smjp_oracle as O
from oracle import philox

pytestmark = pytest.mark.gpu

SHAPE3 = np.array([[1.606, 1.933, 0.865], [1.938, 0.869, 1.751], [1.69, 0.64, 0.696]])


@pytest.fixture(scope="module")
def engine():
    from smjp_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


def test_prior_paths_match_oracle(engine):
    from smjp_b200.chains import DeviceChains
    K, t_end = 3, 10.0
    engine.set_model(SHAPE3, np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, t_end)
    C = 64
    pi0 = np.array([0.2, 0.5, 0.3])
    ch = DeviceChains(engine, C, np.arange(0.1, t_end, 0.1)).init_prior(pi0, seed=42, cap_per_chain=8)
    V, T = ch.paths_host()
    for c in range(C):
        v, w = O.prior_path(SHAPE3, np.ones((K, K)), pi0, t_end, seed=42, chain=c)
        assert np.array_equal(V[c], v)
        assert np.allclose(T[c], w, rtol=1e-13, atol=0)
        assert T[c][0] == 0.0 and T[c][-1] == t_end


def test_prior_honour_scale(engine):
    from smjp_b200.chains import DeviceChains
    rng = np.random.default_rng(1)
    K, t_end = 4, 5.0
    shape = rng.uniform(0.6, 3, (K, K)); scale = rng.uniform(0.5, 2, (K, K))
    engine.set_model(shape, scale, 2.0, O.emission_matrix(K), 1.0, t_end)
    ch = DeviceChains(engine, 16, np.zeros(0)).init_prior(np.ones(K), seed=3, honour_scale=True)
    V, T = ch.paths_host()
    for c in range(16):
        v, w = O.prior_path(shape, scale, np.ones(K), t_end, seed=3, chain=c, honour_scale=True)
        assert np.array_equal(V[c], v) and np.allclose(T[c], w, rtol=1e-13, atol=0)


def test_simulated_observations_match_oracle(engine):
    from smjp_b200.chains import DeviceChains
    K, t_end = 3, 10.0
    emis = O.emission_matrix(K)
    engine.set_model(SHAPE3, np.ones((K, K)), 2.0, emis, 1.0, t_end)
    obs_t = np.arange(0.1, t_end, 0.1)
    ch = DeviceChains(engine, 32, obs_t).init_prior(np.ones(K) / K, seed=9).simulate_observations(seed=10)
    V, T = ch.paths_host()
    y = ch.obs_y.cpu().numpy().reshape(32, len(obs_t))
    for c in range(32):
        assert np.array_equal(y[c], O.simulate_observations(V[c], T[c], obs_t, emis, seed=10, chain=c))


@pytest.mark.parametrize("mode", ["reference", "exact"])
def test_candidates_match_oracle(engine, mode):
    from smjp_b200.engine import candidates_host
    K, t_end, omega = 3, 12.0, 2.0
    engine.set_model(SHAPE3, np.ones((K, K)), omega, O.emission_matrix(K), 1.0, t_end)
    V, T = [], []
    for c in range(40):
        v, w = O.prior_path(SHAPE3, np.ones((K, K)), np.ones(K) / K, t_end, seed=77, chain=c)
        V.append(v); T.append(w)
    Ws = candidates_host(engine, V, T, seed=123, sweep=4, mode=mode)
    for c in range(40):
        ora = O.candidates(V[c], T[c], SHAPE3, np.ones((K, K)), omega, seed=123, sweep=4, chain=c, mode=mode)
        assert len(Ws[c]) == len(ora)
        assert np.allclose(Ws[c], ora, rtol=1e-12, atol=1e-14)
        assert np.all(np.diff(Ws[c]) > 0)


def test_candidates_large_mean_and_omega3(engine):
    """long sojourns -> Poisson means above the 32-chunk; non-unit scales; Omega = 3"""
    from smjp_b200.engine import candidates_host
    rng = np.random.default_rng(8)
    K, t_end, omega = 3, 9.0, 3.0
    shape = rng.uniform(0.6, 3, (K, K)); scale = rng.uniform(0.5, 2, (K, K))
    engine.set_model(shape, scale, omega, O.emission_matrix(K), 1.0, t_end)
    V = [np.array([0, 2, 2]), np.array([1, 1])]
    T = [np.array([0.0, 6.5, 9.0]), np.array([0.0, 9.0])]
    Ws = candidates_host(engine, V, T, seed=5, sweep=0)
    for c in range(2):
        ora = O.candidates(V[c], T[c], shape, scale, omega, seed=5, sweep=0, chain=c)
        assert len(ora) > 60
        assert len(Ws[c]) == len(ora) and np.allclose(Ws[c], ora, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("mode", ["reference", "exact"])
def test_candidates_merge_equal_shape_targets(engine, mode):
    """targets of a row that share a shape are drawn as ONE Poisson process (their superposition): the grids
    equal the oracle's, and the number of thinned events has the mean the separate draws would have"""
    from smjp_b200.engine import candidates_host
    rng = np.random.default_rng(9)
    K, t_end, omega = 4, 6.0, 2.5
    shape = np.array([[1.0, 1.0, 1.0, 1.0], [2.0, 0.7, 2.0, 0.7], [1.3, 1.3, 0.9, 1.1], [0.8, 1.9, 1.5, 0.8]])
    scale = rng.uniform(0.5, 2.0, (K, K))
    engine.set_model(shape, scale, omega, O.emission_matrix(K), 1.0, t_end)
    V, T = [], []
    for c in range(24):
        v, w = O.prior_path(shape, scale, np.ones(K) / K, t_end, seed=31, chain=c, honour_scale=True)
        V.append(v); T.append(w)
    Ws = candidates_host(engine, V, T, seed=17, sweep=2, mode=mode)
    for c in range(24):
        ora = O.candidates(V[c], T[c], shape, scale, omega, seed=17, sweep=2, chain=c, mode=mode)
        assert len(Ws[c]) == len(ora) and np.allclose(Ws[c], ora, rtol=1e-12, atol=1e-14)
    # one long sojourn in state 0 (all four targets merged), many chains: E[#events] = (Omega-1) tau sum 1/lambda
    C = 4000
    Ws = candidates_host(engine, [np.array([0, 0])] * C, [np.array([0.0, t_end])] * C, seed=18, sweep=0, mode=mode)
    cnt = np.array([len(w) - 2 for w in Ws], dtype=np.float64)
    mu = (omega - 1.0) * t_end * (1.0 / scale[0]).sum()
    assert abs(cnt.mean() - mu) < 5 * np.sqrt(mu / C) and abs(cnt.var() - mu) < 0.15 * mu


def test_sweep_host_matches_oracle_composition(engine):
    """sweep = candidates -> FFBS with Philox backward uniforms; compare every stage with the oracle."""
    from smjp_b200.engine import sweep_host
    K, t_end, omega = 3, 8.0, 2.0
    emis = O.emission_matrix(K)
    engine.set_model(SHAPE3, np.ones((K, K)), omega, emis, 1.0, t_end)
    obs_t = np.arange(0.1, t_end, 0.1)
    C = 12
    V, T, Y = [], [], []
    for c in range(C):
        v, w = O.prior_path(SHAPE3, np.ones((K, K)), np.ones(K) / K, t_end, seed=1, chain=c)
        V.append(v); T.append(w)
        Y.append(O.simulate_observations(v, w, obs_t, emis, seed=2, chain=c))
    for sw in range(3):
        r = sweep_host(engine, V, T, [obs_t] * C, Y, seed=99, sweep=sw)
        assert np.all(r["status"] == 0)
        for c in range(C):
            W = O.candidates(V[c], T[c], SHAPE3, np.ones((K, K)), omega, seed=99, sweep=sw, chain=c)
            assert len(W) == len(r["W"][c]) and np.allclose(W, r["W"][c], rtol=1e-12, atol=1e-14)
            n = len(W) - 1
            u = philox.u01(99, philox.TAG_BACKWARD, np.arange(n), 0, sw, c)
            ora = O.ffbs(r["W"][c], obs_t, Y[c], SHAPE3, np.ones((K, K)), omega, emis, 1.0, t_end, u)
            assert np.array_equal(r["V"][c], ora["V"])
            assert np.array_equal(r["T"][c], ora["T"])
            t, j = O.chain_metrics(ora["V"], ora["T"], K)
            assert np.array_equal(r["jumps"][c], j) and np.allclose(r["time_in_state"][c], t, atol=1e-13)
        V, T = r["V"], r["T"]


def test_device_chains_equal_host_sweeps(engine):
    """the device-resident loop (DeviceChains.sweep) and the host-buffer call give identical chains"""
    from smjp_b200.chains import DeviceChains
    from smjp_b200.engine import sweep_host
    K, t_end = 3, 8.0
    emis = O.emission_matrix(K)
    engine.set_model(SHAPE3, np.ones((K, K)), 2.0, emis, 1.0, t_end)
    obs_t = np.arange(0.1, t_end, 0.1)
    C = 50
    ch = DeviceChains(engine, C, obs_t).init_prior(np.ones(K) / K, seed=4).simulate_observations(seed=5)
    V, T = ch.paths_host()
    Y = ch.obs_y.cpu().numpy().reshape(C, -1)
    for sw in range(4):
        ch.sweep(seed=6, sweep=sw)
        r = sweep_host(engine, V, T, [obs_t] * C, list(Y), seed=6, sweep=sw)
        V, T = r["V"], r["T"]
        Vd, Td = ch.paths_host()
        assert all(np.array_equal(a, b) for a, b in zip(V, Vd))
        assert all(np.array_equal(a, b) for a, b in zip(T, Td))
        assert np.array_equal(ch.jumps.cpu().numpy(), r["jumps"])
        assert int(ch.status.abs().sum().item()) == 0


def test_posterior_summaries_ks_vs_oracle_chain(engine):
    """KS (reference acceptance rule, mcmc_utils.py:377-415: reject if D > 1.224 sqrt((n+m)/(nm)))
    between time-in-state / jump counts of GPU chains and of oracle chains run with independent seeds."""
    from scipy import stats
    from smjp_b200.chains import DeviceChains
    K, t_end, omega = 3, 2.0, 2.0
    emis = O.emission_matrix(K)
    engine.set_model(SHAPE3, np.ones((K, K)), omega, emis, 1.0, t_end)
    obs_t = np.arange(0.1, 2.0, 0.1)
    y = np.array([1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 3, 3, 2, 1, 1, 3]) - 1   # experiments.py:217
    C, burn = 600, 12
    ch = DeviceChains(engine, C, obs_t).init_prior(np.ones(K) / K, seed=21).set_observations(y)
    for sw in range(burn):
        ch.sweep(seed=22, sweep=sw)
    g_t = ch.time_in_state.cpu().numpy(); g_j = ch.jumps.cpu().numpy()
    Co = 300  # oracle: independent chains, same number of sweeps, different seed
    o_t, o_j = [], []
    for c in range(Co):
        V, T = O.prior_path(SHAPE3, np.ones((K, K)), np.ones(K) / K, t_end, seed=31, chain=c)
        for sw in range(burn):
            W = O.candidates(V, T, SHAPE3, np.ones((K, K)), omega, seed=32, sweep=sw, chain=c)
            u = philox.u01(32, philox.TAG_BACKWARD, np.arange(len(W) - 1), 0, sw, c)
            r = O.ffbs(W, obs_t, y, SHAPE3, np.ones((K, K)), omega, emis, 1.0, t_end, u)
            V, T = r["V"], r["T"]
        t, j = O.chain_metrics(V, T, K)
        o_t.append(t); o_j.append(j)
    o_t, o_j = np.array(o_t), np.array(o_j)
    ub = 1.224 * np.sqrt((C + Co) / (C * Co))
    for s in range(K):
        assert stats.ks_2samp(g_t[:, s], o_t[:, s]).statistic < ub
        assert stats.ks_2samp(g_j[:, s], o_j[:, s]).statistic < ub


def test_sharded_chains_equal_unsharded_bit_for_bit(engine):
    """'sharded vs unsharded summaries bit-for-bit'.  The Philox streams are keyed by the GLOBAL
    chain index (chain_base + local index), so splitting 64 chains into shards [0,40) and [40,64) -- what
    two ranks would own -- reproduces the single-batch run exactly."""
    from smjp_b200.chains import DeviceChains
    from smjp_b200.engine import Engine
    K, t_end = 3, 6.0
    emis = O.emission_matrix(K)
    obs_t = np.arange(0.1, t_end, 0.1)

    def run(eng, C, base):
        eng.set_model(SHAPE3, np.ones((K, K)), 2.0, emis, 1.0, t_end)
        ch = DeviceChains(eng, C, obs_t, chain_base=base).init_prior(np.ones(K) / K, seed=4).simulate_observations(seed=5)
        ch.init_prior(np.ones(K) / K, seed=6)
        for sw in range(3):
            ch.sweep(seed=7, sweep=sw)
        V, T = ch.paths_host()
        return V, T, ch.time_in_state.cpu().numpy(), ch.jumps.cpu().numpy()

    full = run(engine, 64, 0)
    eng2 = Engine(0)
    a = run(eng2, 40, 0)
    b = run(eng2, 24, 40)
    eng2.close()
    engine.set_option("chain_base", 0)
    V = a[0] + b[0]; T = a[1] + b[1]
    assert all(np.array_equal(x, y) for x, y in zip(V, full[0]))
    assert all(np.array_equal(x, y) for x, y in zip(T, full[1]))
    assert np.array_equal(np.concatenate([a[2], b[2]]), full[2]) and np.array_equal(np.concatenate([a[3], b[3]]), full[3])


@pytest.mark.parametrize("speculate", [1, 2, 0])
def test_packed_sweep_equals_list_api(engine, speculate):
    """PackedSweep (pinned, no per-chain Python work: the e2e path of bench.py) == sweep_host on lists; with the
    normal speculative launches, with forced mis-predictions (2) and with synchronous launches (0)"""
    from smjp_b200.engine import PackedSweep, sweep_host, ragged_offsets
    engine.set_option("speculate", speculate)
    try:
        _packed_sweep_case(engine)
    finally:
        engine.set_option("speculate", 1)


def _packed_sweep_case(engine):
    from smjp_b200.engine import PackedSweep, sweep_host, ragged_offsets
    K, t_end = 3, 8.0
    emis = O.emission_matrix(K)
    engine.set_model(SHAPE3, np.ones((K, K)), 2.0, emis, 1.0, t_end)
    obs_t = np.arange(0.1, t_end, 0.1)
    C = 20
    V, T, Y = [], [], []
    for c in range(C):
        v, w = O.prior_path(SHAPE3, np.ones((K, K)), np.ones(K) / K, t_end, seed=1, chain=c)
        V.append(v); T.append(w); Y.append(O.simulate_observations(v, w, obs_t, emis, seed=2, chain=c))
    ps = PackedSweep(engine, C, ragged_offsets([len(obs_t)] * C), np.tile(obs_t, C), np.concatenate(Y), 64)  # tiny: forces regrowth
    ps.set_paths(V, T)
    for sw in range(6):
        r = sweep_host(engine, V, T, [obs_t] * C, Y, seed=5, sweep=sw)
        ps.run(5, sw)
        for c in range(C):
            lo = ps.w_offs[c]
            assert np.array_equal(ps.W[lo: ps.w_offs[c + 1]], r["W"][c])
            assert np.array_equal(ps.out_v[lo: lo + ps.out_len[c]], r["V"][c])
            assert np.array_equal(ps.out_t[lo: lo + ps.out_len[c]], r["T"][c])
        assert np.array_equal(ps.jumps[: C * K].reshape(C, K), r["jumps"])
        ps.swap()
        V, T = r["V"], r["T"]


@pytest.mark.parametrize("K,precision", [(3, "f64"), (5, "f32")])
def test_pinned_inputs_read_in_place_change_no_result(engine, K, precision):
    """smjp_sweep_host with pinned buffers: the FFBS kernel's bulk copies read each chain's observations straight out of
    the caller's memory and a gather kernel pulls only the path entries in use (option obs_zero_copy, default auto);
    the same sweeps with both switched off (everything uploaded first) give the same bits."""
    from smjp_b200.engine import PackedSweep, ragged_offsets
    rng = np.random.default_rng(K)
    shape = SHAPE3 if K == 3 else rng.uniform(0.6, 3.0, (K, K))
    t_end = 30.0
    emis = O.emission_matrix(K)
    engine.set_model(shape, np.ones((K, K)), 2.0, emis, 1.0, t_end)
    obs_t = np.arange(0.1, t_end, 0.1)
    C = 40
    V, T, Y = [], [], []
    for c in range(C):
        v, w = O.prior_path(shape, np.ones((K, K)), np.ones(K) / K, t_end, seed=1, chain=c)
        V.append(v); T.append(w); Y.append(O.simulate_observations(v, w, obs_t, emis, seed=2, chain=c))
    out = {}
    try:
        for mode in (-1, 0):
            engine.set_option("obs_zero_copy", mode)
            ps = PackedSweep(engine, C, ragged_offsets([len(obs_t)] * C), np.tile(obs_t, C), np.concatenate(Y), 40000)
            ps.set_paths(V, T)
            seen = []
            for sw in range(4):
                ps.run(7, sw, precision=precision)
                seen.append((int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"obs_in_place")),
                             int(engine.lib.smjp_ctx_get_stat(engine.ctx, b"paths_gathered"))))
                assert not np.any(ps.status[:C])
                ps.swap()
            out[mode] = (ps.in_offs[: C + 1].copy(), ps.in_v[: ps.in_offs[C]].copy(), ps.in_t[: ps.in_offs[C]].copy(),
                         ps.in_len[:C].copy(), ps.W[: ps.total_w].copy(), seen)
    finally:
        engine.set_option("obs_zero_copy", -1)
    assert all(s == (0, 0) for s in out[0][5])
    assert all(s[0] == 1 for s in out[-1][5]) and any(s[1] == 1 for s in out[-1][5])
    lens = out[0][3]
    assert np.array_equal(out[0][0], out[-1][0]) and np.array_equal(lens, out[-1][3]) and np.array_equal(out[0][4], out[-1][4])
    for c in range(C):
        lo = out[0][0][c]
        assert np.array_equal(out[0][1][lo: lo + lens[c]], out[-1][1][lo: lo + lens[c]])
        assert np.array_equal(out[0][2][lo: lo + lens[c]], out[-1][2][lo: lo + lens[c]])


def test_two_host_batches_in_flight_equal_the_synchronous_call(engine):
    """smjp_sweep_host_begin / _end: two half-batches on two contexts, one staged while the other computes, give the
    bits of one synchronous smjp_sweep_host call over all chains."""
    from smjp_b200.engine import Engine, PackedSweep, ragged_offsets
    import torch
    K, t_end, C = 3, 30.0, 48
    emis = O.emission_matrix(K)
    model = (SHAPE3, np.ones((K, K)), 2.0, emis, 1.0, t_end)
    engine.set_model(*model)
    obs_t = np.arange(0.1, t_end, 0.1)
    V, T, Y = [], [], []
    for c in range(C):
        v, w = O.prior_path(SHAPE3, np.ones((K, K)), np.ones(K) / K, t_end, seed=1, chain=c)
        V.append(v); T.append(w); Y.append(O.simulate_observations(v, w, obs_t, emis, seed=2, chain=c))
    full = PackedSweep(engine, C, ragged_offsets([len(obs_t)] * C), np.tile(obs_t, C), np.concatenate(Y), 40000)
    full.set_paths(V, T)
    halves = []
    for lo, hi in ((0, 20), (20, C)):
        st = torch.cuda.Stream()
        eh = Engine(0, stream=st.cuda_stream)
        eh.set_model(*model)
        eh.set_option("chain_base", lo)
        ph = PackedSweep(eh, hi - lo, ragged_offsets([len(obs_t)] * (hi - lo)), np.tile(obs_t, hi - lo), np.concatenate(Y[lo:hi]), 20000)
        ph.set_paths(V[lo:hi], T[lo:hi])
        halves.append((lo, hi, st, eh, ph))
    try:
        for _, _, _, _, ph in halves:
            ph.begin(9, 0)
        for sw in range(4):
            full.run(9, sw)
            for lo, hi, _, _, ph in halves:
                ph.end()
                assert not np.any(ph.status[: ph.C])
                for c in range(lo, hi):
                    a, b = full.w_offs[c], ph.w_offs[c - lo]
                    n = full.w_offs[c + 1] - a
                    assert n == ph.w_offs[c - lo + 1] - b and np.array_equal(full.W[a: a + n], ph.W[b: b + n])
                    m = full.out_len[c]
                    assert m == ph.out_len[c - lo]
                    assert np.array_equal(full.out_v[a: a + m], ph.out_v[b: b + m]) and np.array_equal(full.out_t[a: a + m], ph.out_t[b: b + m])
                assert np.array_equal(full.jumps[lo * K: hi * K], ph.jumps[: ph.C * K])
                ph.swap()
                if sw < 3:
                    ph.begin(9, sw + 1)
            full.swap()
    finally:
        for _, _, _, eh, _ in halves:
            eh.close()


@pytest.mark.parametrize("parts", [2, 3])
def test_interleaved_sub_batches_equal_the_single_batch(engine, parts):
    """InterleavedChains: the batch as sub-batches on their own streams and contexts (the next kernel of one fills the
    tail of the other's persistent kernel) -- same chains, same global indices, same bits as one DeviceChains."""
    from smjp_b200.chains import DeviceChains, InterleavedChains
    import torch
    K, T, C = 3, 40.0, 301
    emis = O.emission_matrix(K)
    model = (SHAPE3, np.ones((K, K)), 2.0, emis, 1.0, T)
    obs_t = np.arange(0.1, T, 0.1)
    engine.set_model(*model)
    one = DeviceChains(engine, C, obs_t, chain_base=7).init_prior(np.ones(K) / K, seed=1).simulate_observations(seed=2)
    one.init_prior(np.ones(K) / K, seed=3)
    il = InterleavedChains(0, C, obs_t, model, parts=parts, chain_base=7)
    il.init_prior(np.ones(K) / K, seed=1).simulate_observations(seed=2)
    il.init_prior(np.ones(K) / K, seed=3)
    try:
        for sw in range(5):
            one.sweep(seed=4, sweep=sw)
            il.sweep(seed=4, sweep=sw)
            tis, jm = il.summaries()
            torch.cuda.synchronize()
            assert torch.equal(tis, one.time_in_state) and torch.equal(jm, one.jumps)
        assert int((il.status != 0).sum().item()) == 0
        V1, T1 = one.paths_host()
        V2, T2 = il.paths_host()
        assert all(np.array_equal(a, b) for a, b in zip(V1, V2)) and all(np.array_equal(a, b) for a, b in zip(T1, T2))
    finally:
        il.close()
        engine.set_option("chain_base", 0)


def test_fp32_sweeps_leave_no_failed_chain(engine):
    """1.2e7 fp32 backward draws (4096 chains x 24 sweeps): none may lose its path.  Regression for a draw whose
    target fell between the inclusive scan value of the last thread holding weights and the warp total (a
    shuffle scan over zero-padded lanes is not monotone to the last bit): an empty chunk then 'owned' the draw
    and returned an index one past the row, which repeated a grid point in the path and killed the chain on
    the next sweep.  Every path must stay strictly increasing in time."""
    from smjp_b200.chains import DeviceChains
    K, T, C = 3, 100.0, 4096
    engine.set_model(SHAPE3, np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, T)
    ch = DeviceChains(engine, C, np.arange(0.1, T, 0.1)).init_prior(np.ones(K) / K, seed=1).simulate_observations(seed=2)
    ch.init_prior(np.ones(K) / K, seed=3)
    for sw in range(24):
        ch.sweep(seed=4, sweep=sw, precision="f32")
    assert int((ch.status != 0).sum().item()) == 0
    V, Tt = ch.paths_host()
    assert all(np.all(np.diff(t) > 0) for t in Tt)


def test_live_window_exact_and_pruned_modes(engine):
    """The exact live window skips only grid points whose states are exact zeros (same paths as the oracle, tested
    throughout this suite); here: the counter of processed pairs is below the nominal n(n+1)/2, and the opt-in
    pruned mode (states below 2^-70 of the row sum dropped) processes fewer still and returns the same paths."""
    from smjp_b200.chains import DeviceChains
    K, T, C = 3, 60.0, 256
    engine.set_model(SHAPE3, np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, T)
    obs_t = np.arange(0.1, T, 0.1)
    out = {}
    for wl in (0, -70):
        engine.set_option("window_log2", wl)
        try:
            ch = DeviceChains(engine, C, obs_t).init_prior(np.ones(K) / K, seed=1).simulate_observations(seed=2)
            ch.init_prior(np.ones(K) / K, seed=3)
            for sw in range(3):
                ch.sweep(seed=4, sweep=sw)
            engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_live_pairs")
            ch.sweep(seed=4, sweep=3)
            live = engine.lib.smjp_ctx_get_stat(engine.ctx, b"ffbs_live_pairs")
            nominal = ch.state_steps_last_sweep()[0] / K
            assert int((ch.status != 0).sum().item()) == 0
            out[wl] = (live / nominal, ch.paths_host())
        finally:
            engine.set_option("window_log2", 0)
    assert 0.2 < out[0][0] <= 1.0 and out[-70][0] < out[0][0]
    (Ve, Te), (Vp, Tp) = out[0][1], out[-70][1]
    same = sum(np.array_equal(a, b) and np.array_equal(s, t) for a, b, s, t in zip(Ve, Vp, Te, Tp))
    assert same == C  # dropped mass is below the rounding of the row sums: no draw moves


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_speculative_sweep_launches_change_no_result(engine, precision):
    """smjp_sweep_dev enqueues a sweep's kernels for the grid size predicted from the previous sweep and repeats the
    launch when the prediction was too small (option "speculate").  1: the normal prediction (a few per cent of head
    room); 2: a test setting that predicts 3/4 of the previous size, so that nearly every launch returns at once and
    is repeated; 0: the synchronous sequence.  All three give the same grids and paths."""
    from smjp_b200.chains import DeviceChains
    K, T, C, S = 3, 40.0, 6, 24
    engine.set_model(SHAPE3, np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, T)
    obs_t = np.arange(0.1, T, 0.1)
    out = {}
    for spec in (1, 2, 0):
        engine.set_option("speculate", spec)
        try:
            l0 = engine.lib.smjp_ctx_get_stat(engine.ctx, b"spec_launches")
            m0 = engine.lib.smjp_ctx_get_stat(engine.ctx, b"spec_misses")
            ch = DeviceChains(engine, C, obs_t).init_prior(np.ones(K) / K, seed=1).simulate_observations(seed=2)
            ch.init_prior(np.ones(K) / K, seed=3)
            hist = []
            for sw in range(S):
                ch.sweep(seed=4, sweep=sw, precision=precision)
                hist.append((ch.n_max, ch.total_w))
            assert int((ch.status != 0).sum().item()) == 0
            out[spec] = (ch.paths_host(), hist, engine.lib.smjp_ctx_get_stat(engine.ctx, b"spec_launches") - l0,
                         engine.lib.smjp_ctx_get_stat(engine.ctx, b"spec_misses") - m0)
        finally:
            engine.set_option("speculate", 1)
    (V0, T0), h0, l_off, m_off = out[0]
    assert l_off == 0 and m_off == 0
    for spec in (1, 2):
        (V1, T1), h1, launches, misses = out[spec]
        assert h1 == h0
        assert all(np.array_equal(a, b) and np.array_equal(s, t) for a, b, s, t in zip(V1, V0, T1, T0))
        assert launches == S - 1 and 0 <= misses <= launches
        if spec == 2:
            assert misses >= launches // 2, (launches, misses)


def test_device_summaries_of_posterior_and_prior_chains(engine):
    """the metric history of posterior chains (and of prior draws) stays on the device and is
    reduced there; the batched KS of smjp_b200.summaries equals the reference's host rule
    (mcmc_utils.compute_ks_twosample_time_jump) on the same samples, and ESS / moments are sane."""
    import torch
    from smjp_b200 import mcmc_utils, summaries
    from smjp_b200.chains import DeviceChains
    K, t_end = 3, 2.0
    engine.set_model(SHAPE3, np.ones((K, K)), 2.0, O.emission_matrix(K), 1.0, t_end)
    obs_t = np.arange(0.1, 2.0, 0.1)
    y = np.array([1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 3, 3, 2, 1, 1, 3]) - 1   # experiments.py:217
    C, S = 64, 60
    post = summaries.SummaryHistory(C, K, 16, "cuda:0")
    prior = summaries.SummaryHistory(C, K, 16, "cuda:0")
    ch = DeviceChains(engine, C, obs_t).init_prior(np.ones(K) / K, seed=5).set_observations(y)
    pr = DeviceChains(engine, C, obs_t)
    states = [1, 2, 3]
    for sw in range(S):
        ch.sweep(seed=6, sweep=sw)
        ch.record(post)
        pr.init_prior(np.ones(K) / K, seed=100 + sw)
        V, T = pr.paths_host()
        tt, jj = mcmc_utils.compute_evaluation_chain_metrics({"V": [v + 1 for v in V], "T": T}, states)
        prior.append(torch.tensor(np.array([tt[s] for s in states]).T, device="cuda:0"),
                     torch.tensor(np.array([jj[s] for s in states]).T, device="cuda:0", dtype=torch.int32))
    assert post.n == S and post.buf.is_cuda
    skip = 7
    r = summaries.ks_time_jump(post, prior, skip=skip)
    a = post.samples(skip).reshape(-1, 2 * K).cpu().numpy()
    b = prior.samples(skip).reshape(-1, 2 * K).cpu().numpy()
    host = mcmc_utils.compute_ks_twosample_time_jump({s: a[:, k] for k, s in enumerate(states)}, {s: b[:, k] for k, s in enumerate(states)},
                                                    {s: a[:, K + k] for k, s in enumerate(states)}, {s: b[:, K + k] for k, s in enumerate(states)},
                                                    states, skip=1, verbose=False)
    assert abs(r["ub"] - host["ub"]) < 1e-15
    assert np.allclose(r["times_ks"][:, 0].cpu().numpy(), host["times_ks"][:, 0], atol=1e-12)
    assert np.allclose(r["jumps_ks"][:, 0].cpu().numpy(), host["jumps_ks"][:, 0], atol=1e-12)
    assert np.array_equal(r["reject_times"].cpu().numpy(), host["reject_times"])
    assert bool(r["reject_times"].any())  # 19 observations do move the posterior away from the prior
    ess = post.ess().cpu().numpy()
    assert ess.shape == (C, 2 * K) and np.all(ess > 1.0) and np.all(ess <= S * 4)
    tm, jm = post.moments(burn=10)
    assert abs(float(tm["mean"].sum()) - t_end) < 1e-9  # the times in the K states add up to the horizon


def test_device_chains_regrow_when_the_grid_outgrows_its_buffer(engine):
    """Short prior paths (few entries per chain) leave DeviceChains with a small grid buffer next to a larger,
    recycled path buffer; a model whose sweeps produce long grids must trigger the regrow-and-redo path instead of
    writing past the grid buffer.  Same paths as chains that start with generous buffers."""
    from smjp_b200.chains import DeviceChains
    K, t_end, C = 3, 30.0, 200
    shape = np.full((K, K), 1.0); scale = np.full((K, K), 40.0)   # rare jumps: prior paths of ~2 entries
    engine.set_model(shape, scale, 25.0, O.emission_matrix(K), 1.0, t_end)   # large Omega: many candidates per sojourn
    obs_t = np.arange(0.5, t_end, 0.5)
    out = []
    for generous in (False, True):
        ch = DeviceChains(engine, C, obs_t).init_prior(np.ones(K) / K, seed=1).simulate_observations(seed=2)
        if generous:
            ch.cap = 1 << 20
        for sw in range(4):
            ch.sweep(seed=3, sweep=sw)
        assert int((ch.status != 0).sum().item()) == 0
        assert ch.W.numel() >= ch.total_w
        out.append((ch.paths_host(), ch.total_w))
    (Va, Ta), ta = out[0]
    (Vb, Tb), tb = out[1]
    assert ta == tb and ta > 9 * 3 * C  # the grids really are longer than three times the live path entries
    assert all(np.array_equal(a, b) and np.array_equal(s, t) for a, b, s, t in zip(Va, Vb, Ta, Tb))
