"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, behind oracle/ref_shim.py) in the authoring container.  The reference cannot
travel to the GPU box, so its outputs are committed as small fixtures; re-run with

    python -m oracle.gen_golden

Fixtures (every array is fp64 / int64; state symbols are stored as 0-based indices):
  ffbs_<name>.npz     inputs (W, obs_t, obs_y, shape, scale, omega, power, t_end, uniforms) and the
                      reference's outputs: alpha = exp(log_alphas) [n, K*n] (fast_hmm.py:202-335),
                      samples [n] (fast_hmm.py:104-191 through the RNG seam), V, T
                      (smjp_utils.py:319-330).
  hmm_dense.npz       legacy HiddenMarkovModel (hidden_markov_model.py:213-313, :88-202) driven
                      through adapter callables: log_alphas, output_prob, backward samples.
  metrics.npz         compute_evaluation_chain_metrics (mcmc_utils.py:40-56) on the ffbs paths.
  sampler_stats.npz   draws from the reference's own samplers for KS parity:
                      sample_smjp_event_times (smjp_utils.py:194-210) on a fixed path and
                      sample_smjp_trajectory_prior (:216-254).
"""
import os

import numpy as np

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SHAPE_CFG1 = np.array([[1.606, 1.933, 0.865], [1.938, 0.869, 1.751], [1.69, 0.64, 0.696]])  # experiments.py:172-174
OBS_T_CFG1 = np.arange(0.1, 2.0, 0.1)  # experiments.py:163
OBS_Y_CFG1 = np.array([1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 3, 3, 2, 1, 1, 3])  # experiments.py:217


def ffbs_case(ref, name, states, shape, scale, omega, t_end, obs_t, obs_sym, power, W, uniforms):
    model = ref_shim.build_reference_model(ref, states, shape, scale, omega, t_end, obs_t, obs_sym, power)
    la, samples, V, T, prob = ref_shim.reference_ffbs(ref, model, W, uniforms)
    idx = {s: i for i, s in enumerate(states)}
    np.savez(os.path.join(OUT, "ffbs_%s.npz" % name),
             W=np.asarray(W, dtype=np.float64), obs_t=np.asarray(obs_t, dtype=np.float64),
             obs_y=np.asarray([idx[s] for s in obs_sym], dtype=np.int64),
             shape=np.asarray(shape, dtype=np.float64), scale=np.asarray(scale, dtype=np.float64),
             omega=np.float64(omega), power=np.float64(power), t_end=np.float64(t_end),
             uniforms=np.asarray(uniforms, dtype=np.float64),
             alpha=np.exp(la), samples=samples,
             V=np.asarray([idx[int(round(s))] for s in V], dtype=np.int64), T=T, prob=np.float64(prob))
    print("ffbs_%s: n=%d |V|=%d" % (name, len(W) - 1, len(V)))
    return V, T


def random_grid(rng, t_end, n):
    return np.concatenate([[0.0], np.sort(rng.uniform(0, t_end, n - 1)), [t_end]])


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(20261019)
    paths = []

    # golden vector #1: the configuration of the reference's hand-derived known answer (testing.py:18-31)
    ffbs_case(ref, "golden1", [1, 2], np.array([[1.5, 0.5], [2.0, 0.75]]), np.ones((2, 2)), 2, 1.0,
              np.zeros(0), [], 1.0, [0.0, 1.0 / 3, 2.0 / 3, 1.0], rng.uniform(size=3))

    # default model (experiments.py:154-222) on seeded grids of the sizes the default run produces
    for n in (8, 12, 17):
        V, T = ffbs_case(ref, "cfg1_n%d" % n, [1, 2, 3], SHAPE_CFG1, np.ones((3, 3)), 2, 2.0,
                         OBS_T_CFG1, OBS_Y_CFG1, 1.0, random_grid(rng, 2.0, n), rng.uniform(size=n))
        paths.append((V, T))

    # several draws on one grid: exercises many backward branches
    Wfix = random_grid(rng, 2.0, 9)
    for r in range(4):
        V, T = ffbs_case(ref, "cfg1_rep%d" % r, [1, 2, 3], SHAPE_CFG1, np.ones((3, 3)), 2, 2.0,
                         OBS_T_CFG1, OBS_Y_CFG1, 1.0, Wfix, rng.uniform(size=9))
        paths.append((V, T))

    # non-unit scales, Omega = 3, tempered likelihood
    shape2 = rng.uniform(0.6, 3.0, (3, 3))
    scale2 = rng.uniform(0.5, 2.0, (3, 3))
    ffbs_case(ref, "scaled_omega3", [1, 2, 3], shape2, scale2, 3, 2.0, OBS_T_CFG1, OBS_Y_CFG1, 0.5,
              random_grid(rng, 2.0, 11), rng.uniform(size=11))

    # observations exactly on grid points (counted in both closed intervals) and several
    # observations inside one interval (Q3: summed)
    Wq = np.array([0.0, 0.25, 0.5, 0.75, 1.0, 1.5, 2.0])
    ffbs_case(ref, "obs_on_grid", [1, 2, 3], SHAPE_CFG1, np.ones((3, 3)), 2, 2.0,
              np.array([0.25, 0.3, 0.5, 1.1, 1.2, 1.3, 1.5]), [1, 2, 2, 3, 3, 1, 2], 1.0, Wq, rng.uniform(size=6))

    # no observations at all; 4 states with labels that are not 1..K
    shape4 = rng.uniform(0.6, 3.0, (4, 4))
    ffbs_case(ref, "noobs_k4", [10, 20, 30, 40], shape4, np.ones((4, 4)), 2, 1.5, np.zeros(0), [], 1.0,
              random_grid(rng, 1.5, 7), rng.uniform(size=7))

    # single interval grid (n = 1) and two-interval grid
    ffbs_case(ref, "n1", [1, 2, 3], SHAPE_CFG1, np.ones((3, 3)), 2, 2.0, OBS_T_CFG1, OBS_Y_CFG1, 1.0,
              [0.0, 2.0], rng.uniform(size=1))
    ffbs_case(ref, "n2", [1, 2, 3], SHAPE_CFG1, np.ones((3, 3)), 2, 2.0, OBS_T_CFG1, OBS_Y_CFG1, 1.0,
              [0.0, 0.7, 2.0], rng.uniform(size=2))

    # ---- chain metrics (mcmc_utils.py:40-56) on the sampled paths ----
    agg = {"V": [p[0] for p in paths], "T": [p[1] for p in paths]}
    with ref_shim.quiet():
        st, sj = ref.mcmc.compute_evaluation_chain_metrics(agg, [1, 2, 3])
    L = max(len(p[0]) for p in paths)
    Vp = np.full((len(paths), L), -1, dtype=np.int64)
    Tp = np.full((len(paths), L), np.nan)
    for i, (V, T) in enumerate(paths):
        Vp[i, : len(V)] = np.round(V).astype(np.int64) - 1
        Tp[i, : len(T)] = T
    np.savez(os.path.join(OUT, "metrics.npz"), V=Vp, T=Tp,
             lens=np.array([len(p[0]) for p in paths], dtype=np.int64),
             state_times=np.array([st[s] for s in [1, 2, 3]], dtype=np.float64).T,
             state_jumps=np.array([sj[s] for s in [1, 2, 3]], dtype=np.int64).T)

    # ---- legacy dense HMM (hidden_markov_model.py) through adapter callables ----
    S, n = 4, 12
    P = rng.dirichlet(np.ones(S), size=S)
    E = rng.dirichlet(np.ones(5), size=S)
    pi0 = rng.dirichlet(np.ones(S))
    data = rng.integers(0, 5, size=n)
    uniforms = rng.uniform(size=n)

    class Obs:  # O[time, time_n] -> symbol observed at step time_n
        def __getitem__(self, sl):
            return int(data[int(sl[1])])

    class Pi:
        state_space = np.array([(s, 0.0) for s in range(S)])

        def __call__(self, times):
            return np.log(P)

    class Em:
        def __call__(self, obs):
            return E[:, obs[0]]

    class Init:
        def l(self, s):
            return pi0[s]

    hmm = ref.hmm.HiddenMarkovModel([], emission=Em(), transition=Pi(), data=Obs(),
                                    state_alphabet=list(range(S)), pi_0=Init(),
                                    time_grid=list(range(n)), sample_dimension=1)
    seam = ref_shim.UniformSeam(uniforms[::-1])
    saved = ref.hmm.npr
    ref.hmm.npr = seam
    try:
        with ref_shim.quiet():
            la, prob = hmm.likelihood()
            samples, _ = hmm.backward_sampling(alphas=la)
    finally:
        ref.hmm.npr = saved
    np.savez(os.path.join(OUT, "hmm_dense.npz"), P=P, E=E, pi0=pi0, data=data, uniforms=uniforms,
             log_alphas=np.asarray(la, dtype=np.float64), prob=np.float64(prob),
             samples=np.asarray(samples[0], dtype=np.int64))
    print("hmm_dense: prob=%g" % prob)

    # ---- the reference's own samplers, for KS parity of candidates / prior ----
    np.random.seed(12345)  # the reference never seeds; we do, only to make the fixture reproducible
    model = ref_shim.build_reference_model(ref, [1, 2, 3], SHAPE_CFG1, np.ones((3, 3)), 2, 2.0,
                                           OBS_T_CFG1, OBS_Y_CFG1, 1.0)
    Vfix, Tfix = [1, 3, 2, 2], [0.0, 0.8, 1.3, 2.0]
    counts, first_soj = [], []
    with ref_shim.quiet():
        for _ in range(4000):
            Wd = np.asarray(ref.smjp.sample_smjp_event_times(model.pp_A_hat, Vfix, Tfix, 2.0))
            counts.append(len(Wd))
            first_soj.extend(Wd[(Wd > 0.0) & (Wd < 0.8)].tolist())
        pj, pt = [], []
        for _ in range(4000):
            v, w = ref.smjp.sample_smjp_trajectory_prior(model.hazard_A, model.pi_0, [1, 2, 3], 2.0)
            st, sj = ref.mcmc.compute_evaluation_chain_metrics({"V": [np.asarray(v)], "T": [np.asarray(w)]}, [1, 2, 3])
            pj.append([sj[s][0] for s in [1, 2, 3]])
            pt.append([st[s][0] for s in [1, 2, 3]])
    np.savez(os.path.join(OUT, "sampler_stats.npz"), V=np.array(Vfix) - 1, T=np.array(Tfix),
             shape=SHAPE_CFG1, cand_grid_len=np.array(counts, dtype=np.int64),
             cand_first_sojourn=np.array(first_soj), prior_jumps=np.array(pj, dtype=np.int64),
             prior_times=np.array(pt))
    print("sampler_stats: mean |W| = %.3f, prior mean jumps = %s" % (np.mean(counts), np.mean(pj, axis=0)))


if __name__ == "__main__":
    main()
