"""CPU-only tests: host-side logic of the drop-in layer, ragged packing, and that the C-ABI library
loads and exports every symbol include/smjp_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden

from smjp_b200 import _lib
from smjp_b200.distributions import MultinomialDistribution, WeibullDistribution
from smjp_b200.engine import message_offsets, pack_ragged, ragged_offsets, triangular_to_dense
from smjp_b200 import mcmc_utils, smjp_utils as SU, utils


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "smjp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(smjp_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    assert os.path.isfile(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().smjp_abi_version() == 1


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from smjp_b200.engine import Engine
    with pytest.raises(_lib.EngineError, match="no CUDA device"):
        Engine(0)


def test_ragged_helpers():
    flat, offs = pack_ragged([np.arange(3), np.arange(0), np.arange(5)], np.float64)
    assert offs.tolist() == [0, 3, 3, 8] and flat.tolist() == [0, 1, 2, 0, 1, 2, 3, 4]
    assert ragged_offsets([2, 2]).tolist() == [0, 2, 4]
    assert message_offsets([1, 2, 3], 3).tolist() == [0, 3, 12, 30]
    n, K = 3, 2
    tri = np.arange(K * n * (n + 1) // 2, dtype=np.float64) + 1
    d = triangular_to_dense(tri, n, K).reshape(n, K, n)
    assert d[0, 0, 0] == 1 and d[0, 1, 0] == 2 and d[0, 0, 1] == 0
    assert d[1, 0].tolist() == [3, 4, 0] and d[1, 1].tolist() == [5, 6, 0]
    assert d[2, 1].tolist() == [10, 11, 12]


def test_state_space_enumeration_is_state_major():
    aug = SU.enumerate_state_space([0.0, 0.5, 1.2, 2.0], [1, 2, 3])
    assert aug.shape == (9, 2)
    assert aug[:3].tolist() == [[1, 0.0], [1, 0.5], [1, 1.2]] and aug[3].tolist() == [2, 0.0]  # a = v*n + j


def test_upperbound_scale_and_omega_inference():
    shape = np.array([[1.5, 0.5], [2.0, 0.75]])
    scale = np.array([[1.0, 2.0], [0.5, 1.0]])
    A = SU.smjpHazardFunction([1, 2], shape, scale)
    B = SU.smjpHazardFunction([1, 2], shape, SU.create_upperbound_scale(shape, scale, 3.0), omega=3.0)
    assert abs(SU.infer_omega(A, B) - 3.0) < 1e-12
    # hazards scale exactly by omega
    assert abs(B(0.7, 1, 2) / A(0.7, 1, 2) - 3.0) < 1e-12
    assert abs(B(0.7, 2, None) / A(0.7, 2, None) - 3.0) < 1e-12
    with pytest.raises(ValueError):
        SU.infer_omega(A, SU.smjpHazardFunction([1, 2], shape + 0.1, scale))
    with pytest.raises(ValueError):
        SU.infer_omega(A, SU.smjpHazardFunction([1, 2], shape, scale * np.array([[0.5, 0.4], [0.5, 0.5]])))
    with pytest.raises(ValueError):
        SU.infer_omega(A, SU.smjpHazardFunction([1, 2], shape, SU.create_upperbound_scale(shape, scale, 3.0), omega=2.0))


def test_data_wrapper_and_label_mapping():
    d = SU.sMJPDataWrapper(data=np.array([1, 3, 2, 2]), time=np.array([0.1, 0.2, 0.2, 0.5]))
    assert d[0.2, 0.5].tolist() == [3, 2, 2]  # closed interval on both ends
    assert d[0.3, 0.4].tolist() == []
    with pytest.raises(NotImplementedError):
        d[0.3]
    with pytest.raises(IndexError):
        d[0.1, 0.2, 0.3]
    assert SU.state_index([10, 20, 30], [30.0, 10, 20.0]).tolist() == [2, 0, 1]  # float labels after a sweep (Q14)
    with pytest.raises(ValueError):
        SU.state_index([1, 2], [5])
    t, y = SU.observation_arrays(SU.sMJPDataWrapper(data=[2, 1], time=[0.9, 0.3]), [1, 2])
    assert t.tolist() == [0.3, 0.9] and y.tolist() == [0, 1]


def test_emission_table_and_transition_accessor():
    e = SU.smjpEmission([1, 2, 3], None, 2.0, 1.0)
    E = e.matrix()
    assert np.allclose(E.sum(axis=1), 1) and abs(E[0, 0] - 10 / 12) < 1e-15 and abs(E[0, 1] - 1 / 12) < 1e-15
    assert abs(e.likelihood(0, 1) - 10 / 12) < 1e-15  # observation index 0 under state label 1
    shape = np.array([[1.5, 0.5], [2.0, 0.75]]); scale = np.ones((2, 2))
    A = SU.smjpHazardFunction([1, 2], shape, scale)
    B = SU.smjpHazardFunction([1, 2], shape, SU.create_upperbound_scale(shape, scale, 2), omega=2)
    aug = SU.enumerate_state_space([0.0, 1 / 3, 2 / 3, 1.0], [1, 2])
    pi = SU.smjpTransitionFunction(aug, A, B)
    assert abs(pi(0, 0, 0.0, 1 / 3) - np.log(0.5)) < 1e-12           # thinned: 1 - 1/omega
    assert pi(0, 1, 0.0, 1 / 3) == pytest.approx(np.log(A(1 / 3, 1, 1) / B(1 / 3, 1, None)))  # jump incl. self
    assert pi(1, 0, 1 / 3, 2 / 3) == -np.inf
    with pytest.raises(ValueError):
        pi(0, 0, 0.5, 0.5)
    rows = np.exp([pi(0, c, 0.0, 1 / 3) for c in range(6)])
    assert abs(rows.sum() - 1.0) < 1e-12  # K+1 non-zeros per row, already normalised


def test_chain_metrics_match_reference_fixture():
    g = load_golden("metrics")
    agg = {"V": [g["V"][i, : g["lens"][i]] + 1.0 for i in range(len(g["lens"]))],
           "T": [g["T"][i, : g["lens"][i]] for i in range(len(g["lens"]))]}
    st, sj = mcmc_utils.compute_evaluation_chain_metrics(agg, [1, 2, 3])
    assert np.array_equal(np.array([sj[s] for s in [1, 2, 3]]).T, g["state_jumps"])
    assert np.allclose(np.array([st[s] for s in [1, 2, 3]]).T, g["state_times"], atol=1e-15)


def test_ks_rule_and_pickle_roundtrip(tmp_path, monkeypatch):
    rng = np.random.default_rng(0)
    a = {s: list(rng.normal(size=900)) for s in [1, 2]}
    b = {s: list(rng.normal(size=900)) for s in [1, 2]}
    c = {s: list(rng.normal(loc=1.5, size=900)) for s in [1, 2]}
    r = mcmc_utils.compute_ks_twosample_time_jump(a, b, a, b, [1, 2], verbose=False)
    assert abs(r["ub"] - 1.224 * np.sqrt(60 / 900)) < 1e-12 and not r["reject_times"].any()
    assert mcmc_utils.compute_ks_twosample_time_jump(a, c, a, c, [1, 2], verbose=False)["reject_times"].all()
    monkeypatch.chdir(tmp_path)
    fn = mcmc_utils.save_samples_in_pickle({"V": [[1, 2]]}, 2, "raoteh", "abc", 300)
    assert fn == "results_raoteh_abc_300.pkl"
    agg, u, om = mcmc_utils.load_samples_in_pickle(fn)
    assert agg == {"V": [[1, 2]]} and u == "abc" and om == 2


def test_numerics_helpers_and_distributions():
    assert abs(utils.logsumexp(np.log([0.25, 0.75]))) < 1e-15
    assert utils.np_log(0.0)[0] == -np.inf
    assert utils.isiterable([1]) and not utils.isiterable(3)
    w = WeibullDistribution({"shape": 1.5, "scale": 2.0}, is_hazard_rate=True)
    assert abs(w.l(0.7) - (1.5 / 2.0) * (0.7 / 2.0) ** 0.5) < 1e-14
    dens = WeibullDistribution({"shape": 1.5, "scale": 2.0})
    assert abs(dens.l(0.7) - (1.5 / 2.0) * (0.7 / 2.0) ** 0.5 * np.exp(-(0.7 / 2.0) ** 1.5)) < 1e-14
    import smjp_b200
    smjp_b200.runtime.seed(1)
    s = np.asarray(dens.sample(4000, hold_time=1.0))
    assert s.min() >= 1.0  # conditional on tau > h
    m = MultinomialDistribution({"prob_vector": [0.0, 1.0], "translation": ["a", "b"]})
    assert m.sample(1) == "b" and m.ll(0) == -np.inf and m.l(1) == 1.0


def test_raoteh_argument_checks():
    from smjp_b200.raoteh import inference_error_checking
    with pytest.raises(ValueError):
        inference_error_checking([])
    with pytest.raises(ValueError):
        inference_error_checking(["parameters"])
    inference_error_checking(["trajectory", "parameters"])


def test_columnar_aggregate_round_trip(tmp_path):
    """ragged lists, scalars, dense arrays and the nested 'prior' dict survive the columnar store"""
    from smjp_b200.mcmc_utils import load_aggregate_columnar, save_aggregate_columnar
    rng = np.random.default_rng(0)
    agg = {"W": [rng.uniform(size=int(n)) for n in rng.integers(2, 40, 25)],
           "V": [rng.integers(1, 4, int(n)).astype(np.float64) for n in rng.integers(2, 9, 25)],
           "T": [np.sort(rng.uniform(size=int(n))) for n in rng.integers(2, 9, 25)],
           "prob": [1.0] * 25,
           "time_in_state": rng.uniform(size=(5, 7, 3)),
           "prior": {"W": [np.arange(3.0), np.arange(5.0)], "V": [np.ones(2), np.ones(4)]}}
    fn = save_aggregate_columnar(str(tmp_path / "run"), agg, omega=2.0, uuid_str="abc")
    back, uid, omega = load_aggregate_columnar(fn)
    assert uid == "abc" and omega == 2.0 and back["prob"] == agg["prob"]
    for k in ("W", "V", "T"):
        assert len(back[k]) == len(agg[k]) and all(np.array_equal(a, b) for a, b in zip(back[k], agg[k]))
    assert np.array_equal(back["time_in_state"], agg["time_in_state"])
    assert all(np.array_equal(a, b) for a, b in zip(back["prior"]["W"], agg["prior"]["W"]))


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs on the host cores only (oracle port) and prints ONE JSON line with the
    keys the driver reads"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_batched_summaries_match_numpy_and_scipy():
    """smjp_b200.summaries: batched autocorrelation / ESS / KS / moments == the per-series host
    computations of the drop-in functions (utils.compute_autocorrelation, scipy.stats.ks_2samp, numpy)."""
    import scipy.special
    import scipy.stats as sss
    import torch
    from smjp_b200 import summaries as S
    from smjp_b200.utils import compute_autocorrelation
    rng = np.random.default_rng(0)
    n, C, K = 400, 3, 2
    phi = np.array([0.0, 0.5, 0.9, 0.3]).reshape(1, 1, 4)
    x = np.zeros((n, C, 2 * K))
    e = rng.normal(size=x.shape)
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    xt = torch.from_numpy(x)
    ac = S.autocorrelation(xt).numpy()
    for c in range(C):
        for k in range(2 * K):
            assert np.allclose(ac[:, c, k], compute_autocorrelation(x[:, c, k]), atol=1e-10)
    assert np.all(S.autocorrelation(torch.ones(16, 2, dtype=torch.float64)).numpy() == 0.0)  # constant series
    ess = S.effective_sample_size(torch.from_numpy(x[:, :, :])).numpy()
    # AR(1): ESS/n -> (1 - phi)/(1 + phi); a 400-step series estimates it within a factor ~2
    want = n * (1 - phi.ravel()) / (1 + phi.ravel())
    assert np.all(ess.mean(axis=0) > 0.4 * want) and np.all(ess.mean(axis=0) < 2.5 * want)
    assert ess[:, 2].mean() < ess[:, 1].mean() < ess[:, 0].mean()
    # KS: continuous columns and columns with ties (jump counts are small integers)
    a = np.c_[rng.normal(size=(300, 2)), rng.poisson(3.0, size=(300, 2)).astype(float)]
    b = np.c_[rng.normal(0.3, 1.0, size=(170, 2)), rng.poisson(3.5, size=(170, 2)).astype(float)]
    d, p = S.ks_two_sample(torch.from_numpy(a), torch.from_numpy(b))
    en = np.sqrt(300 * 170 / 470.0)
    for k in range(4):
        r = sss.ks_2samp(a[:, k], b[:, k])
        assert abs(float(d[k]) - r.statistic) < 1e-12
        assert abs(float(p[k]) - scipy.special.kolmogorov(en * r.statistic)) < 1e-9
        assert abs(float(p[k]) - r.pvalue) < 0.05
    # history, thinning, moments
    h = S.SummaryHistory(C, K, 8, "cpu")
    for t in range(n):
        h.append(torch.from_numpy(x[t, :, :K]), torch.from_numpy(np.round(x[t, :, K:]).astype(np.int32)))
    assert h.n == n and h.buf.shape[0] >= n
    tm, jm = h.moments(skip=3, burn=10)
    flat = np.c_[x[10::3, :, :K].reshape(-1, K), np.round(x[10::3, :, K:]).reshape(-1, K)]
    assert np.allclose(tm["mean"].numpy(), flat[:, :K].mean(axis=0)) and np.allclose(jm["var"].numpy(), flat[:, K:].var(axis=0))
    assert np.allclose(tm["median"].numpy(), np.median(flat[:, :K], axis=0))
    assert np.allclose(jm["median"].numpy(), np.median(flat[:, K:], axis=0))
