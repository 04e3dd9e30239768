"""Small numerics helpers with the semantics of the reference's pkg/utils.py (:4-11, :26-31)."""
import numpy as np

from . import runtime


def logsumexp(nd_array):
    """log(sum(exp(x))) without a max shift, exactly as utils.py:4-5 (masked when the sum is 0)."""
    return np.ma.log([np.sum(np.exp(nd_array))])[0]


def sampleDiscrete(probs, N):
    """N categorical draws from `probs` (utils.py:7-8); inverse CDF on uniforms of the package RNG."""
    p = np.asarray(probs, dtype=np.float64)
    cdf = np.cumsum(p / p.sum())
    u = runtime.rng().random(int(N))
    return np.minimum(np.searchsorted(cdf, u, side="right"), len(p) - 1)


def np_log(number):
    """log with log(0) = -inf, returned as a 1-element array like utils.py:10-11."""
    return np.ma.log([number]).filled(-np.inf)


def isiterable(p_object):
    try:
        iter(p_object)
    except TypeError:
        return False
    return True


def compute_autocorrelation(samples):
    """Normalised autocorrelation of a scalar chain (utils.py:36-46), FFT based."""
    x = np.asarray(samples, dtype=np.float64)
    x = x - x.mean()
    n = len(x)
    f = np.fft.rfft(x, 2 * n)
    ac = np.fft.irfft(f * np.conj(f))[:n] / np.arange(n, 0, -1)
    return ac / ac[0] if ac[0] != 0 else ac
