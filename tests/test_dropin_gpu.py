"""GPU tests of the reference-named drop-in layer (smjp_b200.smjp_utils / fast_hmm /
hidden_markov_model / raoteh / pmcmc / experiments) against the reference's golden fixtures."""
import numpy as np
import pytest

from conftest import ffbs_fixture_names, load_golden

pytestmark = pytest.mark.gpu


def build_model(g, labels=None):
    from smjp_b200 import smjp_utils as SU
    from smjp_b200.distributions import MultinomialDistribution
    K = g["shape"].shape[0]
    states = labels or list(range(1, K + 1))
    omega = float(g["omega"])
    tilde = SU.create_upperbound_scale(g["shape"], g["scale"], omega)
    hat = SU.create_upperbound_scale(g["shape"], g["scale"], omega - 1)
    A = SU.smjpHazardFunction(states, g["shape"], g["scale"])
    B = SU.smjpHazardFunction(states, g["shape"], tilde, omega=omega)
    ppB = SU.PoissonProcess(states, None, B, {"shape": g["shape"], "scale": tilde})
    ppA_hat = SU.PoissonProcess(states, None, SU.smjpHazardFunction(states, g["shape"], hat), {"shape": g["shape"], "scale": hat})
    e = SU.smjpEmission(states, ppB, float(g["t_end"]), float(g["power"]))
    data = SU.sMJPDataWrapper(data=[states[i] for i in g["obs_y"]], time=g["obs_t"])
    pi_0 = MultinomialDistribution({"prob_vector": np.ones(K) / K, "translation": states})
    return states, A, B, ppA_hat, ppB, e, data, pi_0


@pytest.mark.parametrize("name", ffbs_fixture_names())
def test_smjp_ffbs_dropin_reproduces_reference(name):
    from smjp_b200 import smjp_utils as SU
    g = load_golden("ffbs_" + name)
    labels = [10, 20, 30, 40] if name == "noobs_k4" else None
    states, A, B, _, _, e, data, _ = build_model(g, labels)
    V, T, prob, r = SU.smjp_ffbs(list(g["W"]), data, states, A, B, e, float(g["t_end"]), "t", uniforms=g["uniforms"],
                                 full_output=True)
    assert prob == 1.0 and V.dtype == np.float64
    assert np.array_equal(V, np.array([float(states[i]) for i in g["V"]]))
    assert np.array_equal(T, g["T"])
    assert np.abs(r["alpha"] - g["alpha"]).max() < 1e-13
    assert e.augmented_state_space.shape == (len(states) * (len(g["W"]) - 1), 2)  # side effect kept


def test_fast_hmm_class_api():
    from smjp_b200 import smjp_utils as SU
    from smjp_b200.fast_hmm import fast_HiddenMarkovModel
    g = load_golden("ffbs_cfg1_n12")
    states, A, B, _, _, e, data, pi_0 = build_model(g)
    aug = SU.enumerate_state_space(list(g["W"]), states)
    e.augmented_state_space = aug
    hmm = fast_HiddenMarkovModel([], emission=e, transition=SU.smjpTransitionFunction(aug, A, B), data=data,
                                 state_alphabet=aug, pi_0=pi_0, time_grid=list(g["W"]), sample_dimension=1, uuid_str="x")
    la, prob = hmm.likelihood()
    assert abs(prob - 1.0) < 1e-12
    ref = np.log(np.where(g["alpha"] > 0, g["alpha"], 1.0)); ref[g["alpha"] == 0] = -np.inf
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(la), fin) and np.abs(la[fin] - ref[fin]).max() < 1e-11
    samples, t_samples = hmm.backward_sampling(alphas=la, uniforms=g["uniforms"])
    assert np.array_equal(samples[0], g["samples"]) and np.array_equal(t_samples[0], aug[g["samples"]])
    edited = la.copy()
    edited[3] = edited[2]           # legal in the reference (fast_hmm.py:41-59 samples from what it is given)
    with pytest.raises(ValueError):  # here: refused loudly, never silently replaced by the engine's own messages
        hmm.backward_sampling(alphas=edited, uniforms=g["uniforms"])
    with pytest.raises(TypeError):
        fast_HiddenMarkovModel([], emission=lambda *a: 0, transition=lambda *a: 0, data=data, state_alphabet=aug,
                               pi_0=pi_0, time_grid=list(g["W"])).likelihood()
    with pytest.raises(ValueError):  # non-increasing grid: smjp_utils.py:398-399
        bad = list(g["W"]); bad[3] = bad[2]
        SU.smjp_ffbs(bad, data, states, A, B, e, float(g["t_end"]), "t")


def test_legacy_hmm_class_api():
    from smjp_b200.distributions import MultinomialDistribution
    from smjp_b200.hidden_markov_model import HiddenMarkovModel, HMMWrapper
    g = load_golden("hmm_dense")
    S = len(g["pi0"])
    hmm = HiddenMarkovModel([], emission=HMMWrapper(g["E"], True), transition=HMMWrapper(np.log(g["P"]), False),
                            data=g["data"], state_alphabet=list(range(S)),
                            pi_0=MultinomialDistribution({"prob_vector": g["pi0"]}), time_grid=np.arange(len(g["data"])),
                            sample_dimension=1)
    la, prob = hmm.likelihood()
    assert np.abs(la - g["log_alphas"]).max() < 1e-10 and abs(prob / float(g["prob"]) - 1) < 1e-10
    s, _ = hmm.backward_sampling(alphas=la, uniforms=g["uniforms"])
    assert np.array_equal(s[0], g["samples"])


def test_event_times_and_prior_dropins():
    import smjp_b200
    from smjp_b200 import smjp_utils as SU
    from oracle import smjp_oracle as O
    g = load_golden("ffbs_cfg1_n8")
    states, A, B, ppA_hat, _, e, data, pi_0 = build_model(g)
    W = SU.sample_smjp_event_times(ppA_hat, [1, 3, 2, 2], [0.0, 0.8, 1.3, 2.0], 2.0, seed=5)
    ora = O.candidates(np.array([0, 2, 1, 1]), np.array([0.0, 0.8, 1.3, 2.0]), g["shape"], g["scale"], 2.0, seed=5, sweep=0, chain=0)
    assert isinstance(W, list) and np.allclose(W, ora, rtol=1e-12)
    ev = ppA_hat.sample([0.2, 1.4], 2, seed=9)
    assert np.all((ev > 0.2) & (ev < 1.4)) and np.all(np.diff(ev) > 0)
    v, w = SU.sample_smjp_trajectory_prior(A, pi_0, states, 2.0, seed=3)
    vo, wo = O.prior_path(g["shape"], g["scale"], np.ones(3) / 3, 2.0, seed=3, chain=0)
    assert v == [states[i] for i in vo] and np.allclose(w, wo, rtol=1e-13) and v[-1] == v[-2] and w[-1] == 2.0
    v2, w2 = SU.sample_smjp_trajectory_prior(A, pi_0, states, 3.0, t_start=1.0, v_0=2, seed=3)
    assert v2[0] == 2 and w2[0] == 1.0 and w2[-1] == 3.0
    smjp_b200.runtime.seed(11)
    a = SU.sample_smjp_event_times(ppA_hat, [1, 1], [0.0, 2.0], 2.0)
    smjp_b200.runtime.seed(11)
    assert a == SU.sample_smjp_event_times(ppA_hat, [1, 1], [0.0, 2.0], 2.0)  # seedable, unlike the reference


def test_raoteh_and_pmcmc_dropins_agree_by_ks():
    """the reference's own acceptance test (experiments.py:297 -> mcmc_utils.py:377-415): Rao-Teh and
    particle MCMC posteriors over the default model must not be distinguishable.  Both samplers are put
    on the SAME target: exact candidate placement and the product likelihood (the reference's FFBS sums
    the emission probabilities of observations that share a grid interval, SURVEY Q3, which its particle
    filter does not -- mirrored by default, switched off here)."""
    import experiments
    from smjp_b200 import mcmc_utils
    from smjp_b200.pmcmc import pmcmc
    from smjp_b200.raoteh import raoteh
    m = experiments.build_experiment_2_model()
    n = 1500
    rt, _, om = raoteh(["trajectory"], n, 300, m["state_space"], m["time_final"], m["emission"], m["data"], m["pi_0"],
                       m["hazard_A"], m["hazard_B"], m["poisson_process_A_hat"], m["poisson_process_B"], "u", m["omega"],
                       m["obs_times"], None, False, seed=1, save=False, candidate_mode="exact", obs_combine="product")
    assert om == 2 and len(rt["V"]) == n and len(rt["W"]) == n and all(p == 1.0 for p in rt["prob"])
    assert all(T[0] == 0.0 and T[-1] == 2.0 and len(V) == len(T) for V, T in zip(rt["V"], rt["T"]))
    # trajectory-only pMCMC: fixed parameters, independent PF proposals accepted by the MH rule
    from smjp_b200.pmcmc import smjp_smc
    from smjp_b200 import runtime
    runtime.seed(2)
    kw = dict(pi_0=m["pi_0"], time_final=2.0, N=200, data=m["data"], A=m["hazard_A"], emission=m["emission"],
              states=m["state_space"], uuid="u", resampler="systematic")
    V, T, ld = smjp_smc(**kw)
    pm = {"V": [], "T": []}
    for i in range(n):
        Vp, Tp, ldp = smjp_smc(**kw)
        if np.log(runtime.rng().uniform()) <= ldp - ld:
            V, T, ld = Vp, Tp, ldp
        pm["V"].append(np.asarray(V, dtype=np.float64)); pm["T"].append(np.asarray(T))
    rt_t, rt_j = mcmc_utils.compute_evaluation_chain_metrics({"V": rt["V"][100:], "T": rt["T"][100:]}, m["state_space"])
    pm_t, pm_j = mcmc_utils.compute_evaluation_chain_metrics({"V": pm["V"][100:], "T": pm["T"][100:]}, m["state_space"])
    r = mcmc_utils.compute_ks_twosample_time_jump(rt_t, pm_t, rt_j, pm_j, m["state_space"], skip=10, verbose=False)
    # both are autocorrelated single chains, so the reference's alpha = 0.05 bound is only indicative here
    # (the independent-sample version of this check is test_posterior_of_batched_raoteh_equals_particle_filter)
    assert (r["times_ks"][:, 0] < 1.6 * r["ub"]).all(), r
    for s in m["state_space"]:
        assert abs(np.mean(rt_t[s]) - np.mean(pm_t[s])) < 0.08
        # the particle filter (like the reference's) does not simulate (t_last_obs, t_final]: ~0.1 fewer jumps
        assert abs(np.mean(rt_j[s]) - np.mean(pm_j[s])) < 0.45


def test_posterior_of_batched_raoteh_equals_particle_filter():
    """independent samples on both sides: the last sweep of 3000 Rao-Teh chains against 3000 independent
    particle-filter paths (N = 4096); two-sample KS on time-in-state per state."""
    import experiments
    from scipy import stats
    from oracle import smjp_oracle as O
    from smjp_b200 import runtime
    from smjp_b200.engine import pf_host
    from smjp_b200.raoteh import raoteh_batch
    from smjp_b200.smjp_utils import emission_arrays, observation_arrays
    m = experiments.build_experiment_2_model()
    C = 3000
    out = raoteh_batch(C, 1, m["state_space"], 2.0, m["emission"], m["data"], m["pi_0"], m["hazard_A"], m["hazard_B"],
                       seed=3, burn_in=50, candidate_mode="exact", obs_combine="product")
    eng = runtime.get_engine()
    eng.set_model(m["hazard_A"].shape_mat, m["hazard_A"].scale_mat, 2.0, emission_arrays(m["emission"]), 1.0, 2.0)
    obs_t, obs_y = observation_arrays(m["data"], m["state_space"])
    for res in ("systematic", "multinomial"):
        # extend=True: simulate (t_last_obs, t_final] too (the reference holds the path flat there)
        r = pf_host(eng, obs_t, obs_y, np.ones(3) / 3, 4096, seed=9, C=C, resampler=res, extend=True)
        tt = np.array([O.chain_metrics(V, T, 3)[0] for V, T in zip(r["V"], r["T"])])
        jj = np.array([O.chain_metrics(V, T, 3)[1] for V, T in zip(r["V"], r["T"])])
        for s in range(3):
            assert stats.ks_2samp(out["time_in_state"][0][:, s], tt[:, s]).pvalue > 1e-3, (res, s)
            assert stats.ks_2samp(out["jumps"][0][:, s], jj[:, s]).pvalue > 1e-3, (res, s)


def test_experiment_2_driver_with_parameter_inference_runs():
    import experiments
    r = experiments.experiment_2(1.0, ["trajectory", "parameters"], number_of_samples=60, number_of_particles=10, seed=4)
    assert set(r) >= {"ub", "times_ks", "jumps_ks"}


def test_raoteh_batch_matches_single_chain_posterior():
    import experiments
    from scipy import stats
    from smjp_b200.raoteh import raoteh_batch
    m = experiments.build_experiment_2_model()
    out = raoteh_batch(512, 4, m["state_space"], 2.0, m["emission"], m["data"], m["pi_0"], m["hazard_A"], m["hazard_B"],
                       seed=3, burn_in=12, candidate_mode="exact")
    assert out["time_in_state"].shape == (4, 512, 3) and np.allclose(out["time_in_state"].sum(axis=2), 2.0)
    out2 = raoteh_batch(512, 1, m["state_space"], 2.0, m["emission"], m["data"], m["pi_0"], m["hazard_A"], m["hazard_B"],
                        seed=99, burn_in=15, candidate_mode="exact")
    for s in range(3):
        assert stats.ks_2samp(out["time_in_state"][-1][:, s], out2["time_in_state"][-1][:, s]).pvalue > 1e-3


def test_raoteh_fused_sweep_equals_the_two_call_path_in_distribution():
    """raoteh() makes ONE host call per iteration (smjp_sweep_host: candidates + FFBS, model uploaded once) when the
    candidate process is (omega - 1) * hazard_A; fused_sweep=False makes the reference's two calls (raoteh.py:58-59).
    Same posterior (time in state), reproducible by seed, grids contain the previous jump times, 'parameters' and
    'prior' inference keep working (the hazards change in place every iteration)."""
    import experiments
    from smjp_b200 import mcmc_utils
    from smjp_b200.raoteh import raoteh
    n = 1200
    out = {}
    for fused in (True, False):
        m = experiments.build_experiment_2_model()
        rt, _, _ = raoteh(["trajectory"], n, 300, m["state_space"], m["time_final"], m["emission"], m["data"], m["pi_0"],
                          m["hazard_A"], m["hazard_B"], m["poisson_process_A_hat"], m["poisson_process_B"], "u", m["omega"],
                          m["obs_times"], None, False, seed=5, save=False, fused_sweep=fused)
        assert len(rt["W"]) == n and all(p == 1.0 for p in rt["prob"])
        assert all(T[0] == 0.0 and T[-1] == 2.0 and len(V) == len(T) and V.dtype == np.float64 for V, T in zip(rt["V"], rt["T"]))
        assert all(set(np.asarray(rt["T"][i])[:-1]).issubset(set(rt["W"][i + 1])) for i in range(0, n - 1, 97))
        out[fused] = mcmc_utils.compute_evaluation_chain_metrics({"V": rt["V"][100:], "T": rt["T"][100:]}, m["state_space"])[0]
    for s in [1, 2, 3]:
        assert abs(np.mean(out[True][s]) - np.mean(out[False][s])) < 0.08
    m = experiments.build_experiment_2_model()
    a1, _, _ = raoteh(["trajectory"], 30, 300, m["state_space"], 2.0, m["emission"], m["data"], m["pi_0"], m["hazard_A"], m["hazard_B"],
                      m["poisson_process_A_hat"], m["poisson_process_B"], "u", 2, m["obs_times"], None, False, seed=9, save=False)
    m = experiments.build_experiment_2_model()
    a2, _, _ = raoteh(["trajectory"], 30, 300, m["state_space"], 2.0, m["emission"], m["data"], m["pi_0"], m["hazard_A"], m["hazard_B"],
                      m["poisson_process_A_hat"], m["poisson_process_B"], "u", 2, m["obs_times"], None, False, seed=9, save=False)
    assert all(np.array_equal(x, y) for x, y in zip(a1["T"], a2["T"]))
    m = experiments.build_experiment_2_model()
    a3, _, _ = raoteh(["trajectory", "parameters", "prior"], 25, 300, m["state_space"], 2.0, m["emission"], m["data"], m["pi_0"],
                      m["hazard_A"], m["hazard_B"], m["poisson_process_A_hat"], m["poisson_process_B"], "u", 2, m["obs_times"], None,
                      False, seed=10, save=False)
    assert len(a3["theta"]) == 25 and len(a3["prior"]["V"]) == 25 and all(T[-1] == 2.0 for T in a3["T"])
