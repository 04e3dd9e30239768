"""Drop-in for the reference's pkg/hidden_markov_model.py: the legacy plain S-state HMM.

Same class / keyword interface (hidden_markov_model.py:10-67).  `transition` and `emission` are
HMMWrapper objects around a log-transition matrix [S,S] (or [n,S,S]) and an emission table [S,Kobs];
`data` holds one symbol per step.  The reference stores unnormalised log messages; the kernel works
scaled and the log scale is added back here, so long chains that underflow in the reference stay finite.
"""
import numpy as np

from . import runtime
from .engine import hmm_dense_host, raise_for_status


class HMMWrapper(object):
    """wrapper for HMM information (hidden_markov_model.py:325-351): `w(obs)[i, j]`"""

    def __init__(self, state_info_array, use_observation):
        self.state_info_array = np.asarray(state_info_array)
        self.observation = None
        self.use_observation = use_observation
        self.state_space = np.c_[np.arange(len(self.state_info_array)), np.zeros(len(self.state_info_array))]

    def __call__(self, observation):
        self.observation = observation
        return self

    def __getitem__(self, a_slice):
        val = self.state_info_array[a_slice]
        if self.use_observation:
            val = val[self.observation]
        self.observation = None
        return val


def _arrays(kwargs):
    pi, e = kwargs["pi"], kwargs["e"]
    if not isinstance(pi, HMMWrapper) and not hasattr(pi, "state_info_array"):
        raise TypeError("the CUDA engine needs HMMWrapper objects (arrays), not Python functors")
    logP = np.asarray(pi.state_info_array, dtype=np.float64)
    emis = np.asarray(e.state_info_array, dtype=np.float64)
    data = np.asarray(kwargs["O"]).astype(np.int32).reshape(-1)
    init = kwargs["init_probs"]
    pi0 = np.asarray(getattr(init, "prob_vector", init), dtype=np.float64)
    if len(kwargs["time_grid"]) != len(data):
        raise ValueError("one observation per grid point is required")
    with np.errstate(divide="ignore"):
        P = np.exp(logP) if kwargs.get("log_transition", True) else logP
    return P, emis, data, pi0


def likelihood_and_decoding_hmm(*args, **kwargs):
    """forward pass (hidden_markov_model.py:213-313): returns (log_alphas unnormalised [n,S],
    output_prob = P(data), log_viterbi, backpath)."""
    opts = {k: kwargs.pop(k) for k in ("precision", "engine") if k in kwargs}
    P, emis, data, pi0 = _arrays(kwargs)
    eng = opts.pop("engine", None) or runtime.get_engine()
    r = hmm_dense_host(eng, P, pi0, emis=emis, y=[data], sample=False, **opts)
    raise_for_status(r["status"])
    with np.errstate(divide="ignore"):
        log_alphas = np.log(r["alpha_hat"][0].astype(np.float64)) + np.cumsum(r["logc"][0])[:, None]
    log_viterbi = np.zeros_like(log_alphas)
    log_viterbi[0] = log_alphas[0]
    return log_alphas, float(np.exp(r["loglik"][0])), log_viterbi, None


def backward_sampling_hmm(*args, **kwargs):
    """s_T ~ alpha_T; s_t ~ alpha_t * pi[:, s_{t+1}] (hidden_markov_model.py:88-202).
    Returns (samples [1, n], t_samples)."""
    opts = {k: kwargs.pop(k) for k in ("precision", "engine", "uniforms", "seed") if k in kwargs}
    P, emis, data, pi0 = _arrays(kwargs)
    eng = opts.pop("engine", None) or runtime.get_engine()
    u = opts.pop("uniforms", None)
    seed = opts.pop("seed", None)
    r = hmm_dense_host(eng, P, pi0, emis=emis, y=[data], uniforms=None if u is None else [u],
                       seed=runtime.next_seed() if seed is None else seed, **opts)
    raise_for_status(r["status"])
    samples = r["states"][0].astype(np.int64)[None, :]
    return samples, [np.asarray(kwargs["pi"].state_space)[samples[0]]]


class HiddenMarkovModel(object):
    """Finite and Discrete Time & State (hidden_markov_model.py:6-67)"""

    def __init__(self, *args, **kwargs):
        self.emission = kwargs["emission"]
        self.transition = kwargs["transition"]
        self.data = kwargs["data"]
        self.state_alphabet = kwargs["state_alphabet"]
        self.pi_0 = kwargs["pi_0"]
        self.time_grid = kwargs["time_grid"]
        self.sample_dimension = kwargs.get("sample_dimension", 1)
        self.log_transition = kwargs.get("log_transition", True)

    def _kw(self):
        return {"num_of_states": len(self.state_alphabet), "time_grid": self.time_grid, "init_probs": self.pi_0,
                "pi": self.transition, "e": self.emission, "O": self.data, "log_transition": self.log_transition}

    def likelihood(self, **opts):
        alphas, output_prob, _, _ = likelihood_and_decoding_hmm([], **self._kw(), **opts)
        return alphas, output_prob

    def decoding(self, **opts):
        _, _, viterbi, backpath = likelihood_and_decoding_hmm([], **self._kw(), **opts)
        return viterbi, backpath

    def learning(self):
        pass

    def backward_sampling(self, num_of_samples=1, alphas=None, **opts):
        return backward_sampling_hmm([], **self._kw(), **opts)
