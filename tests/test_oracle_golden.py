"""The oracle (oracle/smjp_oracle.py) against the reference's golden vectors (CPU only).

Fixtures under tests/golden/ were produced by the UNMODIFIED reference (oracle/gen_golden.py).
"""
import numpy as np
import pytest
from scipy import stats

from conftest import ffbs_fixture_names, load_golden
from oracle import philox, smjp_oracle as O

# golden vector #1: closed-form messages printed by the reference's test (pkg/testing.py:18-108),
# K=2, Omega=2, shapes [[3/2,1/2],[2,3/4]], scale 1, grid [0,1/3,2/3,1], no observations.
GOLDEN1 = np.array([[0.40282252, 0, 0, 0.59717748, 0, 0],
                    [0.21346932, 0.15450294, 0, 0.3430458, 0.28898194, 0],
                    [0.1135967, 0.09052849, 0.22435996, 0.12771929, 0.15586814, 0.28792741]])


def run_oracle(g):
    K = g["shape"].shape[0]
    return O.ffbs(g["W"], g["obs_t"], g["obs_y"], g["shape"], g["scale"], float(g["omega"]),
                  O.emission_matrix(K), float(g["power"]), float(g["t_end"]), g["uniforms"])


def test_golden_vector_1_closed_form():
    g = load_golden("ffbs_golden1")
    r = run_oracle(g)
    assert np.abs(r["alpha"].reshape(3, 6) - GOLDEN1).max() < 6e-9  # constants carry 8 digits
    assert np.abs(g["alpha"] - GOLDEN1).max() < 6e-9  # and the reference HEAD reproduces them


def closed_form_golden1():
    """Independent hand derivation of the same 3x6 matrix from the model definition (A.2)."""
    k = np.array([[1.5, 0.5], [2.0, 0.75]])
    Om = 2.0
    A = lambda v, s, t: k[v, s] * t ** (k[v, s] - 1)
    H = lambda v, t: sum(t ** k[v, s] for s in range(2))
    Av = lambda v, t: A(v, 0, t) + A(v, 1, t)
    t1, t2, t3 = 1 / 3, 2 / 3, 1.0
    a0 = np.zeros((2, 3)); a1 = np.zeros((2, 3)); a2 = np.zeros((2, 3))
    for v in range(2):
        a0[v, 0] = 0.5 * Om * Av(v, t1) * np.exp(-Om * H(v, t1))
    a0 /= a0.sum()
    for v in range(2):
        a1[v, 0] = a0[v, 0] * 0.5 * Om * Av(v, t2) * np.exp(-Om * (H(v, t2) - H(v, t1)))
        a1[v, 1] = Om * Av(v, t1) * np.exp(-Om * H(v, t1)) * sum(
            a0[u, 0] * A(u, v, t1) / (Om * Av(u, t1)) for u in range(2))
    a1 /= a1.sum()
    for v in range(2):  # final interval: no B_v front term
        a2[v, 0] = a1[v, 0] * 0.5 * np.exp(-Om * (H(v, t3) - H(v, t2)))
        a2[v, 1] = a1[v, 1] * 0.5 * np.exp(-Om * (H(v, t2) - H(v, t1)))
        a2[v, 2] = np.exp(-Om * H(v, t1)) * sum(
            a1[u, 0] * A(u, v, t2) / (Om * Av(u, t2)) + a1[u, 1] * A(u, v, t1) / (Om * Av(u, t1))
            for u in range(2))
    a2 /= a2.sum()
    return np.stack([a0.reshape(-1), a1.reshape(-1), a2.reshape(-1)])


def test_golden_vector_1_full_precision():
    g = load_golden("ffbs_golden1")
    cf = closed_form_golden1()
    assert np.abs(cf - GOLDEN1).max() < 6e-9
    assert np.abs(g["alpha"] - cf).max() < 1e-15
    assert np.abs(run_oracle(g)["alpha"].reshape(3, 6) - cf).max() < 1e-15


@pytest.mark.parametrize("name", ffbs_fixture_names())
def test_ffbs_matches_reference(name):
    g = load_golden("ffbs_" + name)
    r = run_oracle(g)
    n = len(g["W"]) - 1
    a = r["alpha"].reshape(n, -1)
    assert a.shape == g["alpha"].shape
    assert np.abs(a - g["alpha"]).max() < 5e-15
    nz = g["alpha"] > 0
    assert np.array_equal(nz, a > 0)
    assert (np.abs(a[nz] - g["alpha"][nz]) / g["alpha"][nz]).max() < 1e-12  # north_star: 1e-6
    assert np.array_equal(r["samples"], g["samples"])  # bit-exact path given the same uniforms
    assert np.array_equal(r["V"], g["V"])
    assert np.array_equal(r["T"], g["T"])
    assert abs(float(g["prob"]) - 1.0) < 1e-12  # Q1: the reference's 'prob' is identically 1
    assert np.all(np.isfinite(r["logZ"]))


def test_chain_metrics_match_reference():
    g = load_golden("metrics")
    for i in range(len(g["lens"])):
        m = int(g["lens"][i])
        times, jumps = O.chain_metrics(g["V"][i, :m], g["T"][i, :m], 3)
        assert np.array_equal(jumps, g["state_jumps"][i])
        assert np.allclose(times, g["state_times"][i], rtol=0, atol=1e-15)


def test_legacy_dense_hmm_matches_reference():
    g = load_golden("hmm_dense")
    E_steps = g["E"][:, g["data"]].T
    la, ah, logc, ll = O.hmm_forward_dense(g["P"], E_steps, g["pi0"])
    assert np.abs(la - g["log_alphas"]).max() < 1e-12
    assert abs(np.exp(ll) - float(g["prob"])) / float(g["prob"]) < 1e-12
    s = O.hmm_backward_dense(ah, g["P"], g["uniforms"])
    assert np.array_equal(s, g["samples"])


def test_collapse_shape1_equals_augmented_filter():
    """Appendix A.5: with every shape == 1 the (v,l) filter marginalised over l is the K-state HMM
    with B = (1-1/Omega) I + diag(1/(Omega A_v)) A and emissions e_i(v) (independent of l)."""
    rng = np.random.default_rng(3)
    K, n, omega, t_end = 4, 14, 2.0, 3.0
    shape = np.ones((K, K)); scale = rng.uniform(0.5, 2.0, (K, K))
    W = np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, n - 1)), [t_end]])
    obs_t = np.sort(rng.uniform(0, t_end, 9)); obs_y = rng.integers(0, K, 9)
    emis = O.emission_matrix(K)
    alpha, logZ = O.forward(W, obs_t, obs_y, shape, scale, omega, emis, 1.0, t_end)
    B = O.collapse_matrix(shape, scale, omega)
    A_v = (1.0 / scale).sum(axis=1)
    E_steps = np.zeros((n, K))
    for i in range(n):
        lx = O.obs_factor(obs_t, obs_y, emis, 1.0, W[i], W[i + 1])
        front = omega * A_v if i < n - 1 else 1.0
        E_steps[i] = lx * front * np.exp(-omega * A_v * (W[i + 1] - W[i]))
    # first step of the (v,l) filter has no transition: alpha_0 = pi0 * e_0
    _, ah, logc, ll = O.hmm_forward_dense(B, E_steps, np.ones(K) / K)
    assert np.abs(alpha.sum(axis=2) - ah).max() < 1e-13
    assert np.abs(logZ - logc).max() < 1e-12


def test_candidate_sampler_distribution_matches_reference():
    """KS: oracle candidates (mode='reference', Philox streams) vs the reference's own
    sample_smjp_event_times draws (fixture sampler_stats.npz)."""
    g = load_golden("sampler_stats")
    lens, first = [], []
    for c in range(3000):
        W = O.candidates(g["V"], g["T"], g["shape"], np.ones((3, 3)), 2.0, seed=77, sweep=0, chain=c)
        assert np.all(np.diff(W) > 0) and W[0] == 0.0 and W[-1] == 2.0
        lens.append(len(W))
        first.extend(W[(W > 0) & (W < 0.8)].tolist())
    assert stats.ks_2samp(lens, g["cand_grid_len"]).pvalue > 1e-3
    assert stats.ks_2samp(first, g["cand_first_sojourn"]).pvalue > 1e-3


def test_prior_sampler_distribution_matches_reference():
    g = load_golden("sampler_stats")
    pj, pt = [], []
    for c in range(3000):
        V, T = O.prior_path(g["shape"], np.ones((3, 3)), np.ones(3) / 3, 2.0, seed=5, chain=c)
        t, j = O.chain_metrics(V, T, 3)
        pj.append(j); pt.append(t)
    pj, pt = np.array(pj), np.array(pt)
    for s in range(3):
        assert stats.ks_2samp(pj[:, s], g["prior_jumps"][:, s]).pvalue > 1e-3
        assert stats.ks_2samp(pt[:, s], g["prior_times"][:, s]).pvalue > 1e-3


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, out in kat:
        assert tuple(int(x) for x in philox.philox4x32_10(*c, *k)) == out
    u = philox.u01(1, philox.TAG_PRIOR, np.arange(100000), 0, 0, 0)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.005


def test_particle_filter_lognormaliser_is_unbiased_for_exact_likelihood():
    """The PF estimate exp(sum Z_d) is unbiased for P(y); check it against a brute-force Monte-Carlo
    estimate from prior paths on a tiny problem (no reference golden exists: SURVEY Q9/Q10)."""
    rng = np.random.default_rng(0)
    K = 2
    shape = np.array([[1.5, 0.5], [2.0, 0.75]]); scale = np.ones((K, K))
    emis = O.emission_matrix(K)
    obs_t = np.array([0.3, 0.6, 0.9]); obs_y = np.array([0, 1, 1])
    ests = [np.exp(O.particle_filter(obs_t, obs_y, shape, scale, emis, np.ones(K) / K, 1.0, 64,
                                     seed=11, chain=c, resampler="systematic")[2]) for c in range(60)]
    # brute force: P(y) = E_prior[prod_d P(y_d | v(t_d))] with scale-honouring prior paths
    tot = []
    for c in range(20000):
        V, T = O.prior_path(shape, scale, np.ones(K) / K, 1.0, seed=99, chain=c, honour_scale=True)
        q = np.searchsorted(T, obs_t, side="right") - 1
        tot.append(np.prod(emis[V[q], obs_y]))
    mc = np.mean(tot); mc_se = np.std(tot) / np.sqrt(len(tot))
    pf = np.mean(ests); pf_se = np.std(ests) / np.sqrt(len(ests))
    assert abs(pf - mc) < 4 * np.hypot(mc_se, pf_se)


def test_collapsed_oracle_is_the_marginal_of_the_augmented_filter():
    """shape == 1 (A.5): dense_ffbs_collapsed's messages equal forward() summed over l, step by step"""
    rng = np.random.default_rng(77)
    K, n, omega, t_end = 4, 25, 2.0, 5.0
    scale = rng.uniform(0.5, 2.0, (K, K)); emis = O.emission_matrix(K)
    W = np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, n - 1)), [t_end]])
    obs_t = np.sort(rng.uniform(0, t_end, 15)); obs_y = rng.integers(0, K, 15)
    alpha, logZ = O.forward(W, obs_t, obs_y, np.ones((K, K)), scale, omega, emis, 1.0, t_end)
    d = O.dense_ffbs_collapsed(W, obs_t, obs_y, scale, omega, emis, 1.0, t_end, rng.uniform(size=n), rng.uniform(size=n))
    assert np.abs(alpha.sum(axis=2) - d["m"]).max() < 1e-12
    assert np.abs(logZ - d["logZ"]).max() < 1e-10
    assert d["T"][0] == 0.0 and d["T"][-1] == t_end and np.all(np.diff(d["T"]) > 0)
    assert set(d["T"]) <= set(W)


def _golden_paths(g):
    offs = np.concatenate([[0], np.cumsum(g["lens"])])
    return [(g["V"][offs[c]: offs[c + 1]], g["T"][offs[c]: offs[c + 1]]) for c in range(len(g["lens"]))]


def test_shape_mle_oracle_matches_reference_fixture():
    """hold_times / weibull_shape_grid restate raoteh.py:136-167: same counts, same grid point, same g"""
    g = load_golden("shape_mle")
    paths = _golden_paths(g)
    K, C = int(g["K"]), len(paths)
    for c in range(C + 1):
        sel = paths if c == C else [paths[c]]
        khat, gv, cnt, _, _ = O.shape_mle_matrix(sel, K)
        assert np.array_equal(cnt, g["count"][c])
        assert np.array_equal(np.isnan(khat), np.isnan(g["khat"][c]))
        ok = ~np.isnan(khat)
        assert np.abs(khat[ok] - g["khat"][c][ok]).max() < 1e-12
        assert np.abs(gv[ok] - g["gval"][c][ok]).max() < 1e-9
