"""GPU parity tests of the parameter step (smjp_shape_mle_dev): holding-time statistics and the reference's
grid-search Weibull shape estimate, against the fixture produced by the unmodified reference and the oracle
(pytest -m gpu)."""
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


def golden_paths(g):
    offs = np.concatenate([[0], np.cumsum(g["lens"])])
    return [g["V"][offs[c]: offs[c + 1]] for c in range(len(g["lens"]))], \
           [g["T"][offs[c]: offs[c + 1]] for c in range(len(g["lens"]))]


def test_shape_mle_matches_reference_fixture(engine):
    from smjp_b200.engine import shape_mle_host
    g = load_golden("shape_mle")
    V, T = golden_paths(g)
    K, C = int(g["K"]), len(V)
    per = shape_mle_host(engine, V, T, K, pooled=False)
    pool = shape_mle_host(engine, V, T, K, pooled=True)
    for c in range(C + 1):
        khat, gv, cnt = (pool["khat"], pool["g"], pool["count"]) if c == C else (per["khat"][c], per["g"][c], per["count"][c])
        assert np.array_equal(cnt, g["count"][c])  # index work: exact
        assert np.array_equal(np.isnan(khat), np.isnan(g["khat"][c]))
        ok = ~np.isnan(khat)
        assert np.abs(khat[ok] - g["khat"][c][ok]).max() < 1e-12  # the same grid point
        assert np.abs(gv[ok] - g["gval"][c][ok]).max() < 1e-9


def test_shape_mle_statistics_and_edge_cases(engine):
    from smjp_b200.engine import shape_mle_host
    rng = np.random.default_rng(5)
    K = 4
    V, T = [], []
    for c in range(9):
        m = int(rng.integers(2, 60))
        V.append(rng.integers(0, K, m))
        T.append(np.concatenate([[0.0], np.cumsum(rng.weibull(1.7, m - 1) * 0.8)]))
    V.append(np.array([2])); T.append(np.array([0.0]))             # a path with no transition at all
    V.append(np.array([1, 1])); T.append(np.array([0.0, 3.0]))     # a single holding time: g < 0 on the whole grid
    r = shape_mle_host(engine, V, T, K, pooled=False)
    for c in range(len(V)):
        khat, gv, cnt, sl, st = O.shape_mle_matrix([(V[c], T[c])], K)
        assert np.array_equal(r["count"][c], cnt)
        assert np.abs(r["sum_log"][c] - sl).max() < 1e-10 and np.abs(r["sum_t"][c] - st).max() < 1e-10
        assert np.array_equal(np.isnan(r["khat"][c]), np.isnan(khat))
        ok = ~np.isnan(khat)
        if ok.any():
            assert np.abs(r["khat"][c][ok] - khat[ok]).max() < 1e-12
    assert r["count"][-2].sum() == 0 and np.all(np.isnan(r["khat"][-2]))
    assert abs(r["khat"][-1][1, 1] - 4.999) < 1e-12                # raoteh.py:153-167 on one sample: the last grid point
    rp = shape_mle_host(engine, V, T, K, pooled=True, grid=0.01, grid_max=3.0)
    khat, gv, cnt, sl, st = O.shape_mle_matrix(list(zip(V, T)), K, grid=0.01, grid_max=3.0)
    assert np.array_equal(rp["count"], cnt) and np.abs(rp["khat"] - khat).max() < 1e-12


def test_dropin_parameter_helpers(engine):
    from smjp_b200 import raoteh as R
    g = load_golden("shape_mle")
    V, T = golden_paths(g)
    labels = [1, 2, 3]
    Vl = np.asarray([labels[v] for v in V[0]], dtype=np.float64)
    h = R.filter_T_by_state(T[0], Vl, 1, 2)
    assert len(h) == g["count"][0][0, 1]
    k, gv = R.esimtate_weibull_shape(h)
    assert abs(k - g["khat"][0][0, 1]) < 1e-12 and abs(gv - g["gval"][0][0, 1]) < 1e-9
    M = R.esimtate_weibull_shape_mat(Vl, T[0], labels)
    ok = ~np.isnan(g["khat"][0])
    assert np.abs(M[ok] - g["khat"][0][ok]).max() < 1e-12
    assert np.all((M[~ok] >= 0.6) & (M[~ok] <= 3.0))


def test_device_chains_shape_mle_recovers_a_shape(engine):
    """a two-state chain that alternates 0 -> 1 -> 0 with Weibull(k) holds: the pooled estimate approaches k"""
    from smjp_b200.chains import DeviceChains
    K, t_end, C = 2, 60.0, 512
    shape = np.array([[9.0, 2.5], [0.8, 9.0]])
    scale = np.array([[50.0, 1.0], [1.0, 50.0]])  # self-jumps practically never win the race
    engine.set_model(shape, scale, 2.0, O.emission_matrix(K), 1.0, t_end)
    ch = DeviceChains(engine, C, np.arange(0.5, t_end, 0.5)).init_prior(np.ones(K) / K, seed=8, honour_scale=True)
    r = ch.shape_mle(pooled=True)
    assert r["count"][0, 1] > 5000 and r["count"][1, 0] > 5000
    assert abs(r["khat"][0, 1] - 2.5) < 0.1 and abs(r["khat"][1, 0] - 0.8) < 0.05
    per = ch.shape_mle(pooled=False)
    assert per["khat"].shape == (C, K, K) and per["count"].sum(axis=0)[0, 1] == r["count"][0, 1]
