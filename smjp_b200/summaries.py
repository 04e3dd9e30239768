"""Posterior summaries of many chains without leaving the device.

The reference turns its aggregate into a report on the host, one Python list per state:
`compute_evaluation_chain_metrics` (mcmc_utils.py:40-56), `compute_metric_summaries` (:58-92), the two-sample
KS acceptance rule `compute_ks_twosample_time_jump` (:377-415) and the unfinished `compute_autocorrelation` /
`compute_ess` (utils.py:36-56).  The FFBS kernels already emit the per-sample metrics (`time_in_state[C,K]`,
`jumps[C,K]`); this module keeps their history on the device and reduces it there, batched over chains and
states, with torch's sort / FFT as the library underneath (bandwidth is trivial next to a sweep: 2K doubles per
chain-sweep).  Every function works on tensors of any device; the drop-in, list-based versions stay in
`smjp_b200.mcmc_utils` / `smjp_b200.utils`.
"""
import math

import torch


class SummaryHistory:
    """[S, C, 2K] history of (time in state, jump counts) per sweep and chain, appended on the device."""

    def __init__(self, C, K, capacity, device):
        self.C, self.K = int(C), int(K)
        self.buf = torch.zeros((int(capacity), self.C, 2 * self.K), dtype=torch.float64, device=device)
        self.n = 0

    def append(self, time_in_state, jumps):
        """one sweep's metrics: time_in_state [C,K] float64, jumps [C,K] int32 (DeviceChains.time_in_state / .jumps)"""
        if self.n == self.buf.shape[0]:  # grow geometrically, on the device
            self.buf = torch.cat([self.buf, torch.zeros_like(self.buf)], dim=0)
        row = self.buf[self.n]
        row[:, : self.K].copy_(time_in_state)
        row[:, self.K:].copy_(jumps)
        self.n += 1
        return self

    def samples(self, skip=1, burn=0):
        """[S', C, 2K]: sweeps burn, burn+skip, ... (the reference thins with [::skip], mcmc_utils.py:381-384)"""
        return self.buf[burn:self.n:skip]

    # ---- compute_metric_summaries (mcmc_utils.py:58-92) over all retained (sweep, chain) samples ----
    def moments(self, skip=1, burn=0):
        x = self.samples(skip, burn).reshape(-1, 2 * self.K)
        K = self.K
        out = {"mean": x.mean(dim=0), "var": x.var(dim=0, unbiased=False), "median": np_median(x)}
        return ({k: v[:K] for k, v in out.items()}, {k: v[K:] for k, v in out.items()})  # (times, jumps)

    def autocorrelation(self, burn=0):
        """[S', C, 2K] normalised autocorrelation along the sweep axis"""
        return autocorrelation(self.buf[burn:self.n])

    def ess(self, threshold=0.01, burn=0):
        """[C, 2K] effective sample sizes of the per-chain metric series"""
        return effective_sample_size(self.buf[burn:self.n], threshold)


def np_median(x):
    """numpy's median along dim 0 (mean of the two middle order statistics for even counts)"""
    s, _ = torch.sort(x, dim=0)
    n = s.shape[0]
    return 0.5 * (s[(n - 1) // 2] + s[n // 2])


def autocorrelation(x):
    """Normalised autocorrelation along dim 0 of x [S, ...], all trailing series at once: FFT of the CENTRED series,
    lag t divided by the S - t products it sums, then by lag 0 (utils.py:36-46 -- which convolves the un-centred
    samples, its centred copy `x` is never used; the definition, and smjp_b200.utils.compute_autocorrelation, centre).
    A constant series gives zeros."""
    S = x.shape[0]
    xc = x - x.mean(dim=0, keepdim=True)
    f = torch.fft.rfft(xc, n=2 * S, dim=0)
    ac = torch.fft.irfft(f * f.conj(), n=2 * S, dim=0)[:S]
    shape = (S,) + (1,) * (x.dim() - 1)
    ac = ac / torch.arange(S, 0, -1, dtype=x.dtype, device=x.device).reshape(shape)
    a0 = ac[:1]
    return torch.where(a0 != 0, ac / torch.where(a0 != 0, a0, torch.ones_like(a0)), torch.zeros_like(ac))


def effective_sample_size(x, threshold=0.01):
    """ESS = S / (1 + 2 sum_{t=1}^{t*-1} rho_t), t* the first lag whose autocorrelation drops below `threshold`
    -- the estimator the reference's unfinished compute_ess(samples, autocorr_threshold=0.01) (utils.py:48-56)
    sets out to compute.  x [S, ...] -> [...]."""
    rho = autocorrelation(x)
    S = rho.shape[0]
    below = rho < threshold
    below[0] = False
    idx = torch.arange(S, device=x.device).reshape((S,) + (1,) * (x.dim() - 1)).expand_as(rho)
    first = torch.where(below, idx, torch.full_like(idx, S)).min(dim=0).values  # t*
    keep = (idx >= 1) & (idx < first.unsqueeze(0))
    tau = 1.0 + 2.0 * (rho * keep).sum(dim=0)
    return S / tau


def ks_two_sample(a, b):
    """Two-sample Kolmogorov-Smirnov statistic for every column: a [n, M], b [m, M] -> (D [M], p [M]).
    D = sup |F_a - F_b| from one sort of the pooled samples (ties are resolved as scipy.stats.ks_2samp does:
    both ECDFs are evaluated AFTER all copies of a value).  p is the asymptotic Kolmogorov tail with the
    effective size nm/(n+m) (scipy switches to it above 10 000 samples and is exact below)."""
    n, m = a.shape[0], b.shape[0]
    pooled = torch.cat([a, b], dim=0)
    srt, order = torch.sort(pooled, dim=0, stable=True)
    from_a = (order < n).to(a.dtype)
    ca = torch.cumsum(from_a, dim=0) / n
    cb = torch.cumsum(1.0 - from_a, dim=0) / m
    last = torch.ones_like(srt, dtype=torch.bool)
    last[:-1] = srt[1:] != srt[:-1]  # evaluate only at the last copy of each distinct value
    d = torch.where(last, (ca - cb).abs(), torch.zeros_like(ca)).max(dim=0).values
    en = math.sqrt(n * m / (n + m))
    lam = (en * d).clamp_min(1e-12)
    j = torch.arange(1, 101, dtype=a.dtype, device=a.device).reshape(-1, 1)
    p = (2.0 * ((-1.0) ** (j - 1)) * torch.exp(-2.0 * (j * lam.reshape(1, -1)) ** 2)).sum(dim=0)
    p = torch.where(lam < 0.2, torch.ones_like(p), p.clamp(0.0, 1.0))  # the series does not converge near 0; p ~ 1 there
    return d, p


def ks_time_jump(post, prior, skip=30, burn=0):
    """compute_ks_twosample_time_jump (mcmc_utils.py:377-415) for two SummaryHistory objects: per state, the KS
    statistic of the thinned posterior samples against the thinned prior samples, for time-in-state and for the
    jump counts, and the reference's acceptance bound 1.224 sqrt((n+m)/(nm)) (alpha = 0.05)."""
    K = post.K
    a = post.samples(skip, burn).reshape(-1, 2 * K)
    b = prior.samples(skip, burn).reshape(-1, 2 * K)
    n, m = a.shape[0], b.shape[0]
    d, p = ks_two_sample(a, b)
    ub = 1.224 * math.sqrt((n + m) / (n * m))
    both = torch.stack([d, p], dim=1)
    return {"ub": ub, "times_ks": both[:K], "jumps_ks": both[K:], "reject_times": d[:K] > ub,
            "reject_jumps": d[K:] > ub, "n": n, "m": m}
