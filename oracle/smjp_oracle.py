"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy, fp64) restatement of the gauenk/smjp hot path.

This is the checker the CUDA engine is compared with.  It is NOT product code: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import it; ``smjp_b200/`` never does (the product fails loudly without its CUDA library).

Pinning status (details in DESIGN.md "Oracle"):
  * forward messages      -- PINNED: golden vector #1 (reference pkg/testing.py:18-108, closed form)
                             and the unmodified reference run behind oracle/ref_shim.py
                             (fixtures tests/golden/ffbs_*.npz, generator oracle/gen_golden.py).
  * backward-sampled path -- PINNED through the RNG seam (ref_shim.UniformSeam) on the same fixtures.
  * path compaction, chain metrics, legacy dense HMM -- PINNED on reference outputs (fixtures).
  * per-step log-normalisers (logZ) -- parity unpinned: the reference normalises every row and
    returns prob == 1.0 (fast_hmm.py:231,321,333); logZ here is the log of exactly those discarded
    normalisers.
  * candidates / prior sampler -- distribution pinned by KS against the reference's samplers;
    stream-level parity unpinned by design (reference uses unseeded numpy.random).
  * particle filter -- parity unpinned: reference smjp_smc has two defects;
    this follows pmcmc.py:224-310 with per-particle copy semantics and per-particle clocks.
  * gamma hazards, systematic resampling, batching -- parity unpinned (absent from the reference).

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
import math

import numpy as np

from . import philox

# ----------------------------------------------------------------------------------------------
# L0 numerics
# ----------------------------------------------------------------------------------------------


def upperbound_scale(shape, scale, c):
    """lambda~ = lambda / c**(1/k)  (smjp_utils.py:180-187).  c = Omega for B, Omega-1 for A_hat."""
    shape = np.asarray(shape, dtype=np.float64)
    return np.asarray(scale, dtype=np.float64) / np.power(float(c), 1.0 / shape)


def weibull_hazard(tau, shape, scale):
    """A_vs(tau) in hazard-rate mode (distributions.py:265-268 via smjp_utils.py:437-457):
    exp(log k - log lam + (k-1)*(log tau - log lam)).  tau: [...], shape/scale: [K,K] -> [K,K,...]."""
    tau = np.asarray(tau, dtype=np.float64)
    k = shape.reshape(shape.shape + (1,) * tau.ndim)
    lam = scale.reshape(scale.shape + (1,) * tau.ndim)
    with np.errstate(divide="ignore"):
        ltau = np.where(tau > 0, np.log(np.where(tau > 0, tau, 1.0)), 0.0)  # np.ma.log(x).filled(0)
    return np.exp(np.log(k) - np.log(lam) + (k - 1.0) * (ltau - np.log(lam)))


def weibull_cumhaz(tau, shape, scale):
    """H_vs(tau) = (tau/lam)**k  (PoissonProcess.weibull_mean, smjp_utils.py:135-145)."""
    tau = np.asarray(tau, dtype=np.float64)
    k = shape.reshape(shape.shape + (1,) * tau.ndim)
    lam = scale.reshape(scale.shape + (1,) * tau.ndim)
    return np.power(tau / lam, k)


def emission_matrix(K, ll_compare=10.0):
    """P(x|v): ll_compare/(K-1+ll_compare) if x==v else 1/(K-1+ll_compare)
    (smjpEmission.d_create, smjp_utils.py:552-557)."""
    E = np.ones((K, K), dtype=np.float64)
    E[np.arange(K), np.arange(K)] = ll_compare
    return E / E.sum(axis=1, keepdims=True)


def inv_cdf(p, u):
    """Categorical draw by inverse CDF on a normalised vector -- the RNG seam's definition
    (ref_shim.UniformSeam.multinomial, standing in for npr.multinomial(1,p) at fast_hmm.py:125,181)."""
    cdf = np.cumsum(p)
    idx = int(np.searchsorted(cdf, u, side="right"))
    last_pos = int(np.nonzero(p > 0)[0][-1])
    return min(idx, last_pos)


# ----------------------------------------------------------------------------------------------
# Forward filter over the augmented (v, l) states  
# ----------------------------------------------------------------------------------------------


def obs_factor(obs_t, obs_y, emis, power, wc, wn, combine="sum"):
    """(sum_{x in [wc,wn]} P(x|v))**power, 1 when the closed interval holds no observation
    (sMJPDataWrapper._slice_2d smjp_utils.py:40-48; compute_likelihood_obs :493-504; :538-542)."""
    K = emis.shape[0]
    sel = np.nonzero((obs_t >= wc) & (obs_t <= wn))[0]
    if len(sel) == 0:
        return np.ones(K)
    if combine == "product":  # not the reference: the proper likelihood of several observations
        lx = np.ones(K)
        for d in sel:
            lx = lx * emis[:, obs_y[d]]
        return lx ** power
    lx = np.zeros(K)
    for d in sel:  # the reference accumulates sample by sample
        lx = lx + emis[:, obs_y[d]]
    return lx ** power


def forward(W, obs_t, obs_y, shape, scale, omega, emis, power, t_end, obs_combine="sum"):
    """Normalised forward messages of likelihood_and_decoding_hmm (fast_hmm.py:202-335) for the
    sMJP transition (smjp_utils.py:382-416) and emission (smjp_utils.py:506-550, :104-133).

    Returns (alpha [n, K, n] with alpha[i, v, j] = P((v, W[j]) | ...) and zeros for dead j > i,
             logZ [n] = log of the row normalisers the reference divides out and discards).
    The reference's dense layout is alpha.reshape(n, K*n): index a = v*n + j (smjp_utils.py:156-178).
    """
    W = np.asarray(W, dtype=np.float64)
    obs_t = np.asarray(obs_t, dtype=np.float64)
    obs_y = np.asarray(obs_y, dtype=np.int64)
    shape = np.asarray(shape, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    n = len(W) - 1
    K = shape.shape[0]
    lam_t = upperbound_scale(shape, scale, omega)  # hazard_B / poisson_process_B parameters
    alpha = np.zeros((n, K, n), dtype=np.float64)
    logZ = np.zeros(n, dtype=np.float64)
    prev = None
    for i in range(n):
        wc, wn = W[i], W[i + 1]
        if not (wn > wc):
            raise ValueError("We can not have time_p >= time_c")  # smjp_utils.py:398-399
        lx = obs_factor(obs_t, obs_y, emis, power, wc, wn, obs_combine)
        l = W[: i + 1]
        t_hold = wn - l
        t_hold_c = wc - l
        # PoissonProcess.likelihood (smjp_utils.py:104-133): sum_s [H~(t_hold) - H~(t_hold_c)]
        log_exp = (weibull_cumhaz(t_hold, shape, lam_t) - weibull_cumhaz(t_hold_c, shape, lam_t)).sum(axis=1)
        if not np.isclose(t_end, wn):  # smjp_utils.py:544-547
            front = weibull_hazard(t_hold, shape, lam_t).sum(axis=1)
            lw = front * np.exp(-log_exp)
        else:
            lw = np.exp(-log_exp)
        e = lx[:, None] * lw  # [K, i+1]
        if i == 0:
            a = e / K  # uniform prior over {(v,0)} (smjp_utils.py:286-289), fast_hmm.py:220-231
        else:
            a = np.zeros((K, i + 1))
            tau = wc - W[:i]  # t_hold = time_c - l_prev  (smjp_utils.py:386)
            A_vs = weibull_hazard(tau, shape, scale)  # [K, K, i]
            A_v = A_vs.sum(axis=1)
            B_v = weibull_hazard(tau, shape, lam_t).sum(axis=1)
            thin = 1.0 - np.exp(np.log(A_v) - np.log(B_v))  # smjp_utils.py:413-415
            a[:, :i] = e[:, :i] * thin * prev
            jump = A_vs / B_v[:, None, :]  # exp(log A_vv' - log B_v)  smjp_utils.py:409-412
            a[:, i] = e[:, i] * np.einsum("vj,vsj->s", prev, jump)
        Z = a.sum()
        logZ[i] = math.log(Z) if Z > 0 else -math.inf
        a = a / Z  # fast_hmm.py:231, :321
        alpha[i, :, : i + 1] = a
        prev = a
    return alpha, logZ


# ----------------------------------------------------------------------------------------------
# Backward sampling (Appendix A.3) and path compaction
# ----------------------------------------------------------------------------------------------


def backward(alpha, W, shape, scale, omega, uniforms):
    """backward_sampling_hmm (fast_hmm.py:104-191) with inverse-CDF draws; uniforms[i] drives the
    draw at grid step i.  Returns samples [n] of dense indices a = v*n + j."""
    W = np.asarray(W, dtype=np.float64)
    n = len(W) - 1
    K = shape.shape[0]
    lam_t = upperbound_scale(shape, scale, omega)
    samples = np.zeros(n, dtype=np.int64)
    p = alpha[n - 1].reshape(-1)
    samples[n - 1] = inv_cdf(p / p.sum(), uniforms[n - 1])
    for i in range(n - 2, -1, -1):
        v_next, j_next = divmod(int(samples[i + 1]), n)
        w = np.zeros((K, n))
        if j_next == i + 1:  # l+ == W[i+1]: a jump happened at W[i+1] (smjp_utils.py:391, 409-412)
            tau = W[i + 1] - W[: i + 1]
            A_to = weibull_hazard(tau, shape, scale)[:, v_next, :]
            B_v = weibull_hazard(tau, shape, lam_t).sum(axis=1)
            w[:, : i + 1] = alpha[i, :, : i + 1] * (A_to / B_v)
        else:  # thinned continuation: one-hot (smjp_utils.py:393, 413-415)
            tau = W[i + 1] - W[j_next]
            A_v = weibull_hazard(np.array([tau]), shape, scale).sum(axis=1)[v_next, 0]
            B_v = weibull_hazard(np.array([tau]), shape, lam_t).sum(axis=1)[v_next, 0]
            w[v_next, j_next] = alpha[i, v_next, j_next] * (1.0 - A_v / B_v)
        p = w.reshape(-1)
        samples[i] = inv_cdf(p / p.sum(), uniforms[i])
    return samples


def compact_path(samples, W, K, t_end):
    """unique (v,l) rows sorted by l, endpoint appended (smjp_utils.py:319-330, :338-348).
    Returns (state indices int64[m], times float64[m])."""
    W = np.asarray(W, dtype=np.float64)
    n = len(W) - 1
    v = samples // n
    j = samples % n
    pairs = sorted(set(zip(j.tolist(), v.tolist())))
    states = [p[1] for p in pairs]
    times = [float(W[p[0]]) for p in pairs]
    if not np.isclose(times[-1], t_end):
        states.append(states[-1])
        times.append(float(t_end))
    return np.asarray(states, dtype=np.int64), np.asarray(times, dtype=np.float64)


def ffbs(W, obs_t, obs_y, shape, scale, omega, emis, power, t_end, uniforms, obs_combine="sum"):
    """smjp_ffbs (smjp_utils.py:261-330) on state indices.  Returns dict."""
    shape = np.asarray(shape, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    alpha, logZ = forward(W, obs_t, obs_y, shape, scale, omega, emis, power, t_end, obs_combine)
    samples = backward(alpha, W, shape, scale, omega, uniforms)
    V, T = compact_path(samples, W, shape.shape[0], t_end)
    return {"alpha": alpha, "logZ": logZ, "samples": samples, "V": V, "T": T}


def chain_metrics(V, T, K):
    """compute_evaluation_chain_metrics (mcmc_utils.py:40-56) for one sample on state indices:
    jumps[s] = #{j : V[j]==s} (counts the first entry and the duplicated end entry),
    times[s] = sum_{j<len-1, V[j]==s} T[j+1]-T[j]."""
    V = np.asarray(V)
    T = np.asarray(T, dtype=np.float64)
    jumps = np.zeros(K, dtype=np.int64)
    times = np.zeros(K, dtype=np.float64)
    for s in range(K):
        jumps[s] = int(np.sum(V == s))
    for idx in range(len(V) - 1):
        times[V[idx]] += T[idx + 1] - T[idx]
    return times, jumps


# ----------------------------------------------------------------------------------------------
# Candidate (thinned-event) sampler and prior sampler, driven by the engine's Philox streams
# ----------------------------------------------------------------------------------------------

POISSON_CHUNK = 32.0


def poisson_inv(mu, u):
    """Poisson(mu) by sequential inversion of one uniform (mu <= POISSON_CHUNK)."""
    p = math.exp(-mu)
    cdf = p
    k = 0
    while u > cdf and k < 1000:
        k += 1
        p *= mu / k
        cdf += p
    return k


def candidates(V, T, shape, scale, omega, seed, sweep, chain, mode="reference"):
    """sample_smjp_event_times (smjp_utils.py:194-210) + PoissonProcess.sample (:61-99).

    Targets of one row with the SAME shape are merged (not the reference, which draws them one by one: the
    superposition of their Poisson processes is the same process, drawn once with the first member's stream).
    Per sojourn q = (V[q], [T[q], T[q+1])) and target (class) s: N ~ Poisson((Omega-1) * H_vs(tau))
    (weibull_mean with lambda_hat, :135-145), then N points.  mode='reference' draws them as the
    reference does -- unit-scale Weibull(k) density truncated to [0,tau] (distributions.py:295 via
    smjp_utils.py:80-84), realised by inverse CDF instead of rejection; mode='exact'
    draws from the density proportional to the hazard, tau * U**(1/k).
    Stream: uniform r of (seed, TAG_CANDIDATES, a = q*K+s, b = sweep, chain); r = 0..m-1 are the
    Poisson chunk uniforms (m = ceil(mu/32)), r = m.. the positions.
    Returns the sorted grid W (float64) = thinned events U T."""
    shape = np.asarray(shape, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    K = shape.shape[0]
    T = np.asarray(T, dtype=np.float64)
    out = []
    for q in range(len(T) - 1):
        v = int(V[q])
        t0, t1 = T[q], T[q + 1]
        tau = t1 - t0
        out.append(t0)
        pts = []
        for s in range(K):
            k = shape[v, s]
            cls = [u for u in range(K) if shape[v, u] == k]  # equal-shape targets superpose into one process
            if cls[0] != s:
                continue  # drawn with the class representative (its first member)
            if len(cls) == 1:
                mu = (omega - 1.0) * (tau / scale[v, s]) ** k
            else:
                mu = (omega - 1.0) * tau ** k * sum(scale[v, u] ** -k for u in cls)
            m = max(1, int(math.ceil(mu / POISSON_CHUNK)))
            a = q * K + s
            N = 0
            for c in range(m):
                N += poisson_inv(mu / m, float(philox.u01(seed, philox.TAG_CANDIDATES, c, a, sweep, chain)))
            for d in range(N):
                u = float(philox.u01(seed, philox.TAG_CANDIDATES, m + d, a, sweep, chain))
                if mode == "reference":
                    F = -math.expm1(-(tau ** k))
                    x = (-math.log1p(-u * F)) ** (1.0 / k)
                else:
                    x = tau * u ** (1.0 / k)
                x = min(x, tau)
                pts.append(t0 + x)
        out.extend(sorted(pts))
    out.append(T[-1])
    return np.asarray(out, dtype=np.float64)


def prior_path(shape, scale, pi0, t_end, seed, chain, honour_scale=False, max_len=1 << 20):
    """sample_smjp_trajectory_prior (smjp_utils.py:216-254): competing Weibull risks, argmin;
    last entry overwritten with (v_prev, t_end) (:251-252).  The reference draws npr.weibull(k),
    i.e. unit scale (distributions.py:295); honour_scale multiplies by lambda.
    Stream (seed, TAG_PRIOR, 0, 0, chain): r=0 picks v0, r = 1 + step*K + s the holding times."""
    shape = np.asarray(shape, dtype=np.float64)
    K = shape.shape[0]
    u0 = float(philox.u01(seed, philox.TAG_PRIOR, 0, 0, 0, chain))
    v = [inv_cdf(np.asarray(pi0, dtype=np.float64) / np.sum(pi0), u0)]
    w = [0.0]
    step = 0
    while w[-1] < t_end and len(v) < max_len:
        us = philox.u01(seed, philox.TAG_PRIOR, 1 + step * K + np.arange(K), 0, 0, chain)
        holds = (-np.log1p(-us)) ** (1.0 / shape[v[-1]])
        if honour_scale:
            holds = holds * scale[v[-1]]
        s = int(np.argmin(holds))
        w.append(w[-1] + float(holds[s]))
        v.append(s)
        step += 1
    v[-1] = v[-2]
    w[-1] = float(t_end)
    return np.asarray(v, dtype=np.int64), np.asarray(w, dtype=np.float64)


def simulate_observations(V, T, obs_t, emis, seed, chain):
    """sample_data_posterior (smjp_utils.py:350-364) + smjpEmission.sample (:559-563): at each
    observation time falling in sojourn [T[q], T[q+1]) emit y ~ emis[V[q], :].
    Stream (seed, TAG_DATA, 0, 0, chain), r = observation index."""
    T = np.asarray(T, dtype=np.float64)
    obs_t = np.asarray(obs_t, dtype=np.float64)
    q = np.clip(np.searchsorted(T, obs_t, side="right") - 1, 0, len(T) - 2)
    us = philox.u01(seed, philox.TAG_DATA, np.arange(len(obs_t)), 0, 0, chain)
    y = np.zeros(len(obs_t), dtype=np.int64)
    for d in range(len(obs_t)):
        y[d] = inv_cdf(emis[int(V[q[d]])], us[d])
    return y


# ----------------------------------------------------------------------------------------------
# Legacy dense K-state HMM (hidden_markov_model.py) and the shape==1 collapse (Appendix A.5)
# ----------------------------------------------------------------------------------------------


def hmm_forward_dense(P, E_steps, pi0):
    """likelihood_and_decoding_hmm (hidden_markov_model.py:213-313): unnormalised forward
    alpha_t[s] = sum_s' alpha_{t-1}[s'] P[s',s] e_t[s], alpha_0 = pi0 * e_0 (:232-238).
    P: [S,S] or [n,S,S] (P[t] used for step t-1 -> t); E_steps: [n, S].
    Computed scaled; returns (log_alpha [n,S] unnormalised as the reference stores it,
    alpha_hat [n,S] row-normalised, logc [n] per-step log scale, loglik = log P(data))."""
    E_steps = np.asarray(E_steps, dtype=np.float64)
    n, S = E_steps.shape
    P = np.asarray(P, dtype=np.float64)
    alpha_hat = np.zeros((n, S))
    logc = np.zeros(n)
    a = np.asarray(pi0, dtype=np.float64) * E_steps[0]
    for t in range(n):
        if t > 0:
            Pt = P if P.ndim == 2 else P[t]
            a = (alpha_hat[t - 1] @ Pt) * E_steps[t]
        c = a.sum()
        logc[t] = math.log(c) if c > 0 else -math.inf
        alpha_hat[t] = a / c
    cum = np.cumsum(logc)
    with np.errstate(divide="ignore"):
        log_alpha = np.log(alpha_hat) + cum[:, None]
    return log_alpha, alpha_hat, logc, float(cum[-1])


def hmm_backward_dense(alpha_hat, P, uniforms):
    """backward_sampling_hmm (hidden_markov_model.py:88-202): s_T ~ alpha_T (:106-107),
    s_t ~ alpha_hat_t * P[:, s_{t+1}] (:140-151, :186).  uniforms[t] drives step t."""
    n, S = alpha_hat.shape
    P = np.asarray(P, dtype=np.float64)
    s = np.zeros(n, dtype=np.int64)
    p = alpha_hat[n - 1]
    s[n - 1] = inv_cdf(p / p.sum(), uniforms[n - 1])
    for t in range(n - 2, -1, -1):
        Pt = P if P.ndim == 2 else P[t + 1]
        p = alpha_hat[t] * Pt[:, s[t + 1]]
        s[t] = inv_cdf(p / p.sum(), uniforms[t])
    return s


def collapse_matrix(shape, scale, omega):
    """Appendix A.5: for shape == 1, B = (1-1/Omega) I + diag(1/(Omega A_v)) A with A_vs = 1/lambda_vs."""
    shape = np.asarray(shape, dtype=np.float64)
    assert np.all(shape == 1.0)
    A = 1.0 / np.asarray(scale, dtype=np.float64)
    K = A.shape[0]
    return (1.0 - 1.0 / omega) * np.eye(K) + A / (omega * A.sum(axis=1, keepdims=True))


def dense_ffbs_collapsed(W, obs_t, obs_y, scale, omega, emis, power, t_end, u_state, u_jump, obs_combine="sum"):
    """FFBS of the shape == 1 sMJP on its K-state collapse (Appendix A.5): the marginal over l of
    forward() above, m_i ~ e_i o (m_{i-1} B) with B = collapse_matrix(), e_i(v) = obs_i(v)
    [Omega A_v]^{W[i+1] != t_end} exp(-Omega A_v (W[i+1]-W[i])) (smjp_utils.py:506-550 with constant hazards).
    Backward: v_i ~ m_i(v) B[v, v_{i+1}] by inverse CDF of u_state[i]; when v_i == v_{i+1} the step was a
    self-jump with probability [A_vv/(Omega A_v)] / B[v,v] (u_jump[i] below it), else the thinned continuation.
    A jump at W[i+1] starts a sojourn there.  Returns dict(m [n,K], logZ [n], V, T) in compact_path() format."""
    W = np.asarray(W, dtype=np.float64)
    obs_t = np.asarray(obs_t, dtype=np.float64)
    obs_y = np.asarray(obs_y, dtype=np.int64)
    scale = np.asarray(scale, dtype=np.float64)
    K = scale.shape[0]
    n = len(W) - 1
    A = 1.0 / scale
    Av = A.sum(axis=1)
    B = collapse_matrix(np.ones_like(scale), scale, omega)
    self_jump = np.diag(A) / (omega * Av)
    pself = self_jump / ((1.0 - 1.0 / omega) + self_jump)
    m = np.zeros((n, K))
    logZ = np.zeros(n)
    for i in range(n):
        wc, wn = W[i], W[i + 1]
        if not (wn > wc):
            raise ValueError("We can not have time_p >= time_c")
        e = obs_factor(obs_t, obs_y, emis, power, wc, wn, obs_combine) * np.exp(-omega * Av * (wn - wc))
        if not np.isclose(t_end, wn):
            e = e * omega * Av
        a = e / K if i == 0 else e * (m[i - 1] @ B)
        Z = a.sum()
        logZ[i] = math.log(Z)
        m[i] = a / Z
    starts = []
    v_next = -1
    for i in range(n - 1, -1, -1):
        w = m[i] if v_next < 0 else m[i] * B[:, v_next]
        v = inv_cdf(w / w.sum(), float(u_state[i]))
        if v_next >= 0 and (v != v_next or float(u_jump[i]) < pself[v]):
            starts.append((float(W[i + 1]), v_next))
        v_next = v
    starts.append((0.0, v_next))
    starts.reverse()
    V = [p[1] for p in starts]
    T = [p[0] for p in starts]
    if not np.isclose(T[-1], t_end):
        V.append(V[-1])
        T.append(float(t_end))
    return {"m": m, "logZ": logZ, "V": np.asarray(V, dtype=np.int64), "T": np.asarray(T, dtype=np.float64)}


# ----------------------------------------------------------------------------------------------
# Parameter step helpers (raoteh.py:119-167)
# ----------------------------------------------------------------------------------------------


def hold_times(V, T, v0, v1):
    """filter_T_by_state (raoteh.py:136-151): T[j+1]-T[j] over consecutive entries (V[j],V[j+1]) == (v0,v1)."""
    V = np.asarray(V)
    T = np.asarray(T, dtype=np.float64)
    sel = np.nonzero((V[:-1] == v0) & (V[1:] == v1))[0]
    return T[sel + 1] - T[sel]


def weibull_shape_grid(h, grid=0.001, grid_max=5.0):
    """esimtate_weibull_shape (raoteh.py:153-167): the first k of arange(grid, grid_max, grid) minimising
    |sum T^k ln T / sum T^k - 1/k - mean ln T|; returns (k, g(k)) -- (-1, inf) if every g is NaN."""
    h = np.asarray(h, dtype=np.float64)
    ks = np.arange(grid, grid_max, grid)
    lt = np.log(h)
    tk = np.power(h[None, :], ks[:, None])
    g = (tk * lt).sum(axis=1) / tk.sum(axis=1) - 1.0 / ks - lt.mean()
    a = np.abs(g)
    if np.all(np.isnan(a)):
        return -1.0, math.inf
    m = int(np.nanargmin(a))
    return float(ks[m]), float(g[m])


def shape_mle_matrix(paths, K, grid=0.001, grid_max=5.0):
    """esimtate_weibull_shape_mat (raoteh.py:119-134) on the holding times of one or several paths pooled in
    order; NaN where a transition never occurs (the reference draws U(0.6,3) there).
    Returns (khat [K,K], g [K,K], count [K,K], sum_log [K,K], sum_t [K,K])."""
    khat = np.full((K, K), np.nan)
    g = np.full((K, K), np.nan)
    cnt = np.zeros((K, K), dtype=np.int64)
    sl = np.zeros((K, K))
    st = np.zeros((K, K))
    for a in range(K):
        for b in range(K):
            h = np.concatenate([hold_times(V, T, a, b) for V, T in paths]) if len(paths) else np.zeros(0)
            cnt[a, b] = len(h)
            if len(h):
                khat[a, b], g[a, b] = weibull_shape_grid(h, grid, grid_max)
                sl[a, b] = np.log(h).sum()
                st[a, b] = h.sum()
    return khat, g, cnt, sl, st


# ----------------------------------------------------------------------------------------------
# Bootstrap particle filter (Appendix A.6; pmcmc.py:224-310 with the Q9/Q10 defects removed)
# ----------------------------------------------------------------------------------------------


def cond_weibull(u, hold, k, lam):
    """WeibullDistribution.sample(hold_time=h) (distributions.py:296-304): u' ~ U(F(h),1),
    tau = lam*(-log(1-u'))**(1/k).  Written in the algebraically identical, cancellation-free form
    -log(1-u') = (h/lam)**k - log(1-u)."""
    return lam * np.power(np.power(hold / lam, k) - np.log1p(-u), 1.0 / k)


def cond_gamma(u, hold, a, theta):
    """Gamma(shape a, scale theta) holding time given tau > hold, by inverting the survival function
    Q(a, tau/theta) = (1-u) Q(a, hold/theta).  Not in the reference (its hazards are Weibull only,
    smjp_utils.py:419-423; distributions.py:179-217 has the gamma density); parity unpinned."""
    from scipy import special
    return theta * special.gammainccinv(a, (1.0 - u) * special.gammaincc(a, hold / theta))


def logsumexp(x):
    """log sum exp; the reference's (utils.py:4-5) has no max shift -- identical wherever finite."""
    m = np.max(x)
    if not np.isfinite(m):
        return m
    return m + math.log(np.sum(np.exp(x - m)))


def resample_ancestors(w, N, resampler, seed, d, chain):
    """Ancestor indices.  'multinomial' = sampleDiscrete (utils.py:7-8, pmcmc.py:302) by inverse CDF
    with N uniforms; 'systematic' (new; not in the reference) uses one uniform: positions (u+p)/N.
    Stream (seed, TAG_PF_RESAMPLE, 0, d, chain), r = particle (multinomial) or 0 (systematic)."""
    cdf = np.cumsum(w)
    cdf = cdf / cdf[-1]
    if resampler == "systematic":
        u = float(philox.u01(seed, philox.TAG_PF_RESAMPLE, 0, 0, d, chain))
        pos = (u + np.arange(N)) / N
    else:
        pos = philox.u01(seed, philox.TAG_PF_RESAMPLE, np.arange(N), 0, d, chain)
    return np.minimum(np.searchsorted(cdf, pos, side="right"), N - 1).astype(np.int64)


def particle_filter(obs_t, obs_y, shape, scale, emis, pi0, t_end, N, seed, chain,
                    resampler="systematic", max_jumps=64, extend=False, family="weibull"):
    """smjp_smc (pmcmc.py:224-310).  Per observation d: every particle (v, l) is propagated by
    competing conditional-Weibull risks (:261-289) until its next jump would pass obs_t[d]; weight
    log P(y_d | v) (:279-281); Z_d = logsumexp(logw) - log N (:294-297); resample (:302-304).
    Propagation stream (seed, TAG_PF_PROPAGATE, particle, d, chain), r = iteration*K + s; the
    initial states use d = 0xffffffff, r = 0.
    Returns (path_states, path_times, loglik, Z [D]) -- path of one particle after the last resample
    (index 0 for multinomial as the reference, a uniformly drawn index for systematic) plus
    (v_last, t_end) (:307-309)."""
    shape = np.asarray(shape, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    K = shape.shape[0]
    D = len(obs_y)
    t = np.concatenate([[0.0], np.asarray(obs_t, dtype=np.float64)])
    u0 = philox.u01(seed, philox.TAG_PF_PROPAGATE, 0, np.arange(N), 0xFFFFFFFF, chain)
    cdf0 = np.cumsum(np.asarray(pi0, dtype=np.float64) / np.sum(pi0))
    v = np.minimum(np.searchsorted(cdf0, u0, side="right"), K - 1).astype(np.int64)
    l = np.zeros(N)
    paths = [([int(v[p])], [0.0]) for p in range(N)]
    Z = np.zeros(D)
    last_anc = np.arange(N)
    for d in range(D):
        t_obs = t[d + 1]
        t_cur = np.full(N, t[d])
        for p in range(N):
            vp, lp, tc = int(v[p]), float(l[p]), float(t_cur[p])
            pv, pt = paths[p]
            for it in range(max_jumps):
                us = philox.u01(seed, philox.TAG_PF_PROPAGATE, it * K + np.arange(K), p, d, chain)
                taus = (cond_gamma if family == "gamma" else cond_weibull)(us, tc - lp, shape[vp], scale[vp])
                s = int(np.argmin(taus))
                w_next = lp + float(taus[s])
                if w_next >= t_obs:
                    break
                tc = w_next
                lp = w_next
                vp = s
                pv.append(vp)
                pt.append(lp)
            v[p], l[p] = vp, lp
        logw = np.log(emis[v, int(obs_y[d])])
        Zd = logsumexp(logw)
        w = np.exp(logw - Zd)
        Z[d] = Zd - math.log(N)
        anc = resample_ancestors(w, N, resampler, seed, d, chain)
        last_anc = anc
        v = v[anc]
        l = l[anc]
        paths = [(list(paths[a][0]), list(paths[a][1])) for a in anc]  # copy semantics (fixes Q9)
    k = 0  # multinomial offspring are i.i.d.: the reference's index 0 (pmcmc.py:307) is a fair draw
    if resampler == "systematic":  # offspring are ordered by ancestor: draw the index uniformly
        k = min(int(float(philox.u01(seed, philox.TAG_PF_RESAMPLE, 1, 0, D, chain)) * N), N - 1)
    ps, pt = paths[k]
    if extend:  # simulate (t_last_obs, t_end] as well (the reference extends the path flat, pmcmc.py:307-309)
        vp, lp, tc = int(v[k]), float(l[k]), float(t[D])
        slot = int(last_anc[k]) if D > 0 else 0
        for it in range(max_jumps):
            us = philox.u01(seed, philox.TAG_PF_PROPAGATE, it * K + np.arange(K), slot, D, chain)
            taus = (cond_gamma if family == "gamma" else cond_weibull)(us, tc - lp, shape[vp], scale[vp])
            s = int(np.argmin(taus))
            if lp + float(taus[s]) >= t_end:
                break
            tc = lp = lp + float(taus[s])
            vp = s
            ps.append(vp)
            pt.append(lp)
    return np.asarray(ps + [ps[-1]], dtype=np.int64), np.asarray(pt + [float(t_end)]), float(np.sum(Z)), Z
