"""Distribution objects the sMJP model is assembled from (interface of the reference's
pkg/distributions.py: `.l` likelihood, `.ll` log-likelihood, `.s` sample; params by keyword dict).
Only the two on the hot path are provided: Weibull (hazard / density / conditional sampling,
distributions.py:233-307) and the categorical the reference calls Multinomial (:310-378)."""
import numpy as np

from . import runtime


class _Params(dict):
    """keyword parameters with attribute access (the reference uses easydict)"""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class Distribution:
    name = "Distribution"

    def __init__(self, params, is_hazard_rate=False):
        self.params = _Params(params)
        self.is_hazard_rate = is_hazard_rate
        self.prior = None

    def set_prior(self, prior):
        self.prior = prior

    def set_parameters(self, params):
        self.params.update(params)

    def log_likelihood(self, x, params=None, **kw):
        raise NotImplementedError

    def likelihood(self, x, params=None, **kw):
        return np.exp(self.log_likelihood(x, params=params, **kw))

    def sample(self, n=1, params=None, **kw):
        raise NotImplementedError

    # the reference's shorthands
    def ll(self, x, params=None, **kw):
        return self.log_likelihood(x, params=params, **kw)

    def nll(self, x, params=None, **kw):
        return -self.log_likelihood(x, params=params, **kw)

    def l(self, x, params=None, **kw):  # noqa: E743
        return self.likelihood(x, params=params, **kw)

    def s(self, n=1, params=None, **kw):
        return self.sample(n=n, params=params, **kw)

    def _p(self, params):
        return self.params if params is None else _Params({**self.params, **params})

    def __str__(self):
        return "Distribution type [%s] with params: %s" % (self.name, dict(self.params))


class WeibullDistribution(Distribution):
    """shape k, scale lambda.  is_hazard_rate=True makes `.l(t)` the hazard (k/lam)(t/lam)^(k-1)
    (distributions.py:265-268) instead of the density."""
    name = "Weibull"

    def __init__(self, params, params_keywords=None, params_transforms=None, is_hazard_rate=False, **kw):
        super().__init__(params, is_hazard_rate)

    @property
    def shape(self):
        return self.params["shape"]

    @property
    def scale(self):
        return self.params["scale"]

    def log_likelihood(self, x, params=None, **kw):
        p = self._p(params)
        k, lam = p["shape"], p["scale"]
        xs = np.atleast_1d(np.asarray(x, dtype=np.float64))
        with np.errstate(divide="ignore"):
            lx = np.where(xs > 0, np.log(np.where(xs > 0, xs, 1.0)), 0.0)  # log 0 -> 0 as np.ma.log(.).filled(0)
        val = np.log(k) - np.log(lam) + (k - 1.0) * (lx - np.log(lam))
        if not self.is_hazard_rate:
            val = val - (xs / lam) ** k
        val = np.where(xs < 0, 0.0, val)  # negative holding times contribute nothing (:264)
        return float(val[0]) if np.ndim(x) == 0 else float(val.sum())

    def sample(self, n=1, params=None, hold_time=None, **kw):
        """hold_time=None: npr.weibull(k) -- UNIT scale, as the reference (:295).
        hold_time=h: tau | tau > h with the scale honoured (:296-304)."""
        p = self._p(params)
        k, lam = p["shape"], p["scale"]
        u = runtime.rng().random(int(n))
        if hold_time is None:
            out = (-np.log1p(-u)) ** (1.0 / k)
        else:
            out = lam * ((hold_time / lam) ** k - np.log1p(-u)) ** (1.0 / k)
        return float(out[0]) if n == 1 else out


class MultinomialDistribution(Distribution):
    """categorical over len(prob_vector) indices, optionally translated to labels (:310-378)"""
    name = "Multinomial"

    def __init__(self, params, params_keywords=None, params_transforms=None, **kw):
        super().__init__(params)
        self.params.setdefault("translation", None)

    @property
    def prob_vector(self):
        return self.params["prob_vector"]

    @property
    def translation(self):
        return self.params["translation"]

    def log_likelihood(self, x, params=None, **kw):
        p = np.asarray(self._p(params)["prob_vector"], dtype=np.float64)
        if np.ndim(x) != 0:
            raise NotImplementedError("Multinomial likelihood of a sequence is not defined in the reference either")
        with np.errstate(divide="ignore"):
            return float(np.log(p[int(x)]))

    def sample(self, n=1, params=None, **kw):
        pr = self._p(params)
        p = np.asarray(pr["prob_vector"], dtype=np.float64)
        cdf = np.cumsum(p / p.sum())
        idx = np.minimum(np.searchsorted(cdf, runtime.rng().random(int(n)), side="right"), len(p) - 1)
        tr = pr.get("translation")
        out = [tr[i] for i in idx] if tr is not None else [int(i) for i in idx]
        return out[0] if n == 1 else out


class Gamma(Distribution):
    """Gamma(shape, rate) density (distributions.py:179-217, with gammaln(shape) where the reference
    has gammaln(rate), :203).  Not on any reference path; kept for the gamma-hazard extension."""
    name = "Gamma"

    def __init__(self, params, params_keywords=None, params_transforms=None, **kw):
        super().__init__(params)

    def log_likelihood(self, x, params=None, **kw):
        from scipy.special import gammaln
        p = self._p(params)
        xs = np.atleast_1d(np.asarray(x, dtype=np.float64))
        return float(np.sum(p["shape"] * np.log(p["rate"]) - gammaln(p["shape"]) + (p["shape"] - 1) * np.log(xs) - p["rate"] * xs))

    def sample(self, n=1, params=None, **kw):
        p = self._p(params)
        return runtime.rng().gamma(p["shape"], 1.0 / p["rate"], int(n))
