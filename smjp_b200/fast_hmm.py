"""Drop-in for the reference's pkg/fast_hmm.py: the HMM engine over the augmented (v,l) states.

fast_HiddenMarkovModel takes the same keyword arguments (fast_hmm.py:11-19).  `emission` must be an
smjpEmission and `transition` an smjpTransitionFunction (or the reference's own instances): the
engine evaluates them in-kernel from their parameters, it cannot call arbitrary Python functors --
which is precisely the O(n (Kn)^2) callback loop this engine removes (fast_hmm.py:296, :343-354).
"""
import numpy as np

from . import runtime
from .engine import raise_for_status, triangular_to_dense
from .smjp_utils import load_model, observation_arrays


def _engine_inputs(kwargs):
    e, pi = kwargs["e"], kwargs["pi"]
    for obj, need in ((e, "final_grid_time"), (pi, "hazard_A")):
        if not hasattr(obj, need):
            raise TypeError("the CUDA engine needs an smjpEmission / smjpTransitionFunction (got %s)" % type(obj).__name__)
    states = list(e.state_space)
    W = np.asarray(kwargs["time_grid"], dtype=np.float64)
    return e, pi, states, W


def _run(kwargs, want_messages, uniforms=None, seed=None, precision="f64", engine=None):
    e, pi, states, W = _engine_inputs(kwargs)
    eng = engine or runtime.get_engine()
    load_model(eng, states, pi.hazard_A, pi.hazard_B, e, e.final_grid_time)
    obs_t, obs_y = observation_arrays(kwargs["O"], states)
    r = eng.ffbs_host([W], [obs_t], [obs_y], uniforms=None if uniforms is None else [np.asarray(uniforms, dtype=np.float64)],
                      seed=runtime.next_seed() if seed is None else seed, precision=precision, want_messages=want_messages)
    raise_for_status(r["status"])
    return r, len(W) - 1, len(states)


def likelihood_and_decoding_hmm(*args, **kwargs):
    """forward filter (fast_hmm.py:202-335).  Returns (log_alphas [n, K*n], output_prob, log_viterbi,
    backpath): rows are normalised, so output_prob == 1.0 as in the reference; the
    discarded log-normalisers are available as likelihood_and_decoding_hmm.last_logZ."""
    opts = {k: kwargs.pop(k) for k in ("precision", "engine") if k in kwargs}
    r, n, K = _run(kwargs, True, **opts)
    with np.errstate(divide="ignore"):
        log_alphas = np.log(triangular_to_dense(r["messages"][0], n, K).astype(np.float64))
    likelihood_and_decoding_hmm.last_logZ = r["logZ"][0]
    log_viterbi = np.zeros_like(log_alphas)
    log_viterbi[0] = log_alphas[0]  # the reference fills row 0 only: Viterbi is not implemented there
    return log_alphas, float(np.sum(np.exp(log_alphas[-1]))), log_viterbi, None


def backward_sampling_hmm(*args, **kwargs):
    """backward sampling (fast_hmm.py:104-191).  Returns (samples [1, n] dense indices a = v*n + j,
    t_samples = [aug_state_space[samples]]).  The forward messages are recomputed on the device (the
    fused kernel costs less than uploading `alphas`); when `alphas` is given it is checked against them and a
    mismatch raises ValueError.  `uniforms=` / `seed=` make the draw reproducible."""
    opts = {k: kwargs.pop(k) for k in ("precision", "engine", "uniforms", "seed") if k in kwargs}
    alphas = kwargs.get("alphas")
    r, n, K = _run(kwargs, alphas is not None, **opts)
    if alphas is not None:
        # The reference samples from whatever `alphas` it is handed (fast_hmm.py:41-59).  The engine recomputes the
        # forward pass inside the same kernel, so messages that are NOT the ones likelihood() returns for this model
        # (a caller that edited them) cannot be honoured: refuse loudly instead of silently sampling something else.
        mine = triangular_to_dense(r["messages"][0], n, K).astype(np.float64)
        given = np.exp(np.asarray(alphas, dtype=np.float64))
        tol = 1e-9 if opts.get("precision", "f64") in ("f64", "fp64", "float64") else 1e-4
        if given.shape != mine.shape or not np.allclose(given, mine, rtol=tol, atol=tol * 1e-3):
            raise ValueError("backward_sampling(alphas=...): the messages differ from this model's forward pass; the CUDA "
                             "engine samples from its own forward messages and cannot use edited ones")
    samples = r["aug_idx"][0].astype(np.int64)[None, :]
    aug = np.asarray(kwargs["pi"].state_space)
    return samples, [aug[samples[0]]]


class fast_HiddenMarkovModel(object):
    """Finite and Discrete Time & State (fast_hmm.py:7-71)"""

    def __init__(self, *args, **kwargs):
        self.emission = kwargs["emission"]
        self.transition = kwargs["transition"]
        self.data = kwargs["data"]
        self.state_alphabet = kwargs["state_alphabet"]
        self.pi_0 = kwargs["pi_0"]  # ignored by the reference as well: uniform over (v, 0)
        self.time_grid = kwargs["time_grid"]
        self.sample_dimension = kwargs.get("sample_dimension", 1)
        self.uuid_str = kwargs.get("uuid_str")

    def _kw(self):
        return {"num_of_states": len(self.state_alphabet), "time_grid": self.time_grid, "init_probs": self.pi_0,
                "pi": self.transition, "e": self.emission, "O": self.data, "uuid": self.uuid_str}

    def likelihood(self, **opts):
        alphas, output_prob, _, _ = likelihood_and_decoding_hmm([], **self._kw(), **opts)
        return alphas, output_prob

    def decoding(self, **opts):
        _, _, viterbi, backpath = likelihood_and_decoding_hmm([], **self._kw(), **opts)
        return viterbi, backpath

    def learning(self):
        pass

    def backward_sampling(self, num_of_samples=1, alphas=None, **opts):
        kw = self._kw()
        kw.update({"num_of_samples": num_of_samples, "sample_dimension": self.sample_dimension, "alphas": alphas})
        return backward_sampling_hmm([], **kw, **opts)
