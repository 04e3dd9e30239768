"""Drop-in for the reference's pkg/smjp_utils.py: the sMJP model objects and the three sweep kernels
(candidate sampler, FFBS, prior sampler) under their reference names and positional signatures.

The objects only HOLD parameters; all batched arithmetic runs in the CUDA engine.  The engine reads
  smjpHazardFunction.{state_space, shape_mat, scale_mat, omega}      (smjp_utils.py:425-429)
  PoissonProcess.mean_function_params{'shape','scale'}               (:55-59, :147-151)
  smjpEmission.{state_space, likelihood_power, final_grid_time,
                how_likely_is_current_state_compared_to_others}       (:484-491)
  sMJPDataWrapper.{data, time}                                        (:21-23)
by attribute, so the reference's own instances are accepted as well (duck typing).
Every function takes keyword-only extras (`seed=`, `uniforms=`, `precision=`, `engine=`) that the
reference does not have; without them it behaves like the reference call with a fresh seed.
"""
import numpy as np

from . import runtime
from .distributions import MultinomialDistribution, WeibullDistribution
from .engine import candidates_host, raise_for_status, triangular_to_dense
from .utils import isiterable

Weibull = WeibullDistribution
Multinomial = MultinomialDistribution


# ------------------------------------------------------------------------------------------------
# model objects
# ------------------------------------------------------------------------------------------------
class sMJPDataWrapper(object):
    """observations `data[d]` at times `time[d]`; `wrapper[t0, t1]` = symbols with t0 <= time <= t1
    (closed interval, smjp_utils.py:40-48)."""

    def __init__(self, data, time):
        self.data = data
        self.time = time

    def __getitem__(self, a_slice):
        if not isiterable(a_slice):
            raise NotImplementedError("Why are you using only a 1d slice for observations?")
        if len(a_slice) != 2:
            raise IndexError("only a matrix; two arguments please.")
        t0, t1 = a_slice
        t = np.asarray(self.time, dtype=np.float64)
        return np.asarray(self.data)[(t >= t0) & (t <= t1)]

    def __str__(self):
        return str(np.c_[self.data, self.time])


class smjpHazardFunction(object):
    """Weibull hazards A_vs(tau) = (k/lam)(tau/lam)^(k-1) ("hazard function is weibull for now")."""

    def __init__(self, state_space, shape_mat, scale_mat, omega=None):
        self.state_space = state_space
        self.shape_mat = shape_mat
        self.scale_mat = scale_mat
        self.omega = omega

    def update_parameters(self, parameters):
        if "shape" in parameters:
            self.shape_mat = parameters["shape"]
        if "scale" in parameters:
            self.scale_mat = parameters["scale"]

    def _idx(self, state):
        assert state in self.state_space, "must have the state in state space"
        return list(self.state_space).index(state)

    def h_create(self, state_curr, state_next):
        i, j = self._idx(state_curr), self._idx(state_next)
        return Weibull({"shape": self.shape_mat[i][j], "scale": self.scale_mat[i][j]}, is_hazard_rate=True)

    def __call__(self, observation, s_curr, s_next):
        """hazard of leaving s_curr for s_next after holding `observation`; s_next=None sums over
        all targets including s_curr itself (smjp_utils.py:447-457)."""
        targets = self.state_space if s_next is None else [s_next]
        return sum(self.h_create(s_curr, s).l(observation) for s in targets)

    def sample(self, current_state, next_state, hold_time, n=1):
        if hold_time is None:
            return self.h_create(current_state, next_state).sample(n)
        if next_state is None:
            rates = np.array([self.h_create(current_state, s).l(hold_time) for s in self.state_space])
            idx = Multinomial({"prob_vector": rates / rates.sum()}).sample(n)
            idx = [idx] if n == 1 else idx
            return [self.state_space[i] for i in idx]
        return self.h_create(current_state, next_state).sample(n, hold_time=hold_time)


class PoissonProcess(object):
    """inhomogeneous Poisson process with Weibull cumulative intensity (t/lam)^k per target state"""

    def __init__(self, state_space, mean_function, hazard_function, mean_function_params=None):
        self.state_space = state_space
        self.mean_function = self.weibull_mean
        self.mean_function_params = mean_function_params
        self.hazard_function = hazard_function

    def weibull_mean(self, t_start, t_end, current_state, state):
        i = list(self.state_space).index(current_state)
        j = list(self.state_space).index(state)
        k = self.mean_function_params["shape"][i][j]
        lam = self.mean_function_params["scale"][i][j]
        return (t_end / lam) ** k - (t_start / lam) ** k

    def update_parameters(self, parameters):
        for key in ("shape", "scale"):
            if key in parameters:
                self.mean_function_params[key] = parameters[key]

    def sample(self, interval, current_state, seed=None, mode="reference", engine=None):
        """events of the process on `interval` for a sojourn in `current_state`, sorted and offset
        (smjp_utils.py:61-99) -- one-sojourn call of the candidates kernel."""
        t0, t1 = float(interval[0]), float(interval[1])
        if t1 - t0 < 0:
            raise ValueError("negative sojourn length %g" % (t1 - t0))  # the reference exit()s here (:69-72)
        W = sample_smjp_event_times(self, [current_state, current_state], [t0, t1], t1, seed=seed, mode=mode,
                                    engine=engine)
        return np.asarray(W[1:-1], dtype=np.float64)

    def likelihood(self, interval, current_state, is_poisson_event=True):
        """[sum_s hazard(tau)] * exp(-sum_s mean(interval)) (smjp_utils.py:104-133); scalar accessor."""
        ex = sum(self.weibull_mean(interval[0], interval[1], current_state, s) for s in self.state_space)
        if not is_poisson_event:
            return np.exp(-ex)
        return sum(self.hazard_function(interval[1], current_state, s) for s in self.state_space) * np.exp(-ex)

    def l(self, *args, **kwargs):  # noqa: E743
        return self.likelihood(*args, **kwargs)


class smjpEmission(object):
    """P(x | v) = c/(K-1+c) if x == v else 1/(K-1+c), c = how_likely_is_current_state_compared_to_others
    (smjp_utils.py:552-557), tempered by likelihood_power."""

    def __init__(self, state_space, p_w, time_final, likelihood_power, ll_compare=10):
        self.state_space = state_space
        self.p_x = self.likelihood
        self.p_w = p_w
        self.augmented_state_space = None
        self.likelihood_power = likelihood_power
        self.how_likely_is_current_state_compared_to_others = ll_compare
        self.final_grid_time = time_final

    def d_create(self, state_curr):
        p = np.ones(len(self.state_space))
        p[list(self.state_space).index(state_curr)] = self.how_likely_is_current_state_compared_to_others
        return Multinomial({"prob_vector": p / p.sum(), "translation": self.state_space})

    def sample(self, s_curr, n=1):
        return self.d_create(s_curr).s(n=n)

    def likelihood(self, observation, s_curr):
        return self.d_create(s_curr).l(observation)

    def matrix(self):
        """emis[v, x] for the engine"""
        K = len(self.state_space)
        E = np.ones((K, K))
        E[np.arange(K), np.arange(K)] = self.how_likely_is_current_state_compared_to_others
        return E / E.sum(axis=1, keepdims=True)


class smjpTransitionFunction(object):
    """log pi((v,l) -> (v',l'); w_p, w_c) over the augmented states (smjp_utils.py:382-416).  The
    engine evaluates these entries in-kernel; this callable is the scalar accessor."""

    def __init__(self, augmented_state_space, hazard_A, hazard_B):
        self.state_space = augmented_state_space
        self.augmented_state_space = augmented_state_space
        self.hazard_A = hazard_A
        self.hazard_B = hazard_B

    def __call__(self, s_prev_idx, s_curr_idx, time_p, time_c):
        v_p, l_p = self.augmented_state_space[s_prev_idx]
        v_c, l_c = self.augmented_state_space[s_curr_idx]
        if time_p >= time_c:
            raise ValueError("We can not have time_p >= time_c")
        t_hold = time_c - l_p
        jump = (l_c == time_c) and (l_p <= time_p)
        thin = (l_c == l_p) and (v_c == v_p) and (l_p <= time_p)
        if l_p > l_c or t_hold <= 0 or not (jump or thin):
            return -np.inf
        with np.errstate(divide="ignore"):
            lA = np.log(self.hazard_A(t_hold, v_p, None))
            lB = np.log(self.hazard_B(t_hold, v_p, None))
            assert lA - lB <= 0, "the ratio of hazard functions should be <= 1"
            if jump:
                return float(np.log(self.hazard_A(t_hold, v_p, v_c)) - lB)
            return float(np.log(1.0 - np.exp(lA - lB)))


# ------------------------------------------------------------------------------------------------
# helpers shared by the drop-in modules
# ------------------------------------------------------------------------------------------------
def enumerate_state_space(grid, states):
    """augmented states = states x grid[:-1], state-major: row v*n + j = (states[v], grid[j])
    (smjp_utils.py:156-178)."""
    g = list(grid[:-1])
    if 0 not in g:
        g = [0] + g
    mesh = np.meshgrid(g, states)
    return np.c_[mesh[1].ravel(), mesh[0].ravel()]


def create_upperbound_scale(shape_mat, scale_mat, constant):
    """lambda / constant**(1/k): multiplies hazard and cumulative hazard by `constant` (:180-187)"""
    return np.asarray(scale_mat, dtype=np.float64) / np.power(float(constant), 1.0 / np.asarray(shape_mat, dtype=np.float64))


def include_endpoint_in_thinned_samples(thinned_samples, t_end):
    if np.isclose(thinned_samples[-1, 1], t_end):
        return thinned_samples
    return np.r_[thinned_samples, np.array([[thinned_samples[-1, 0], t_end]])]


def state_index(state_space, labels):
    """labels -> 0-based indices.  The reference uses list.index; labels come back from FFBS as
    float64 (smjp_utils.py:328), so compare numerically when possible."""
    states = list(state_space)
    out = np.empty(len(labels), dtype=np.int32)
    for i, s in enumerate(labels):
        try:
            out[i] = states.index(s)
        except ValueError:
            hit = [k for k, x in enumerate(states) if np.isclose(float(x), float(s))]
            if not hit:
                raise ValueError("state %r is not in the state space %r" % (s, states))
            out[i] = hit[0]
    return out


def infer_omega(hazard_A, hazard_B):
    """B must be omega * A (experiments.py:177: scale_B = scale_A / omega**(1/k)); returns omega."""
    kA, kB = np.asarray(hazard_A.shape_mat, dtype=np.float64), np.asarray(hazard_B.shape_mat, dtype=np.float64)
    lA, lB = np.asarray(hazard_A.scale_mat, dtype=np.float64), np.asarray(hazard_B.scale_mat, dtype=np.float64)
    if kA.shape != kB.shape or not np.allclose(kA, kB, rtol=1e-12, atol=0):
        raise ValueError("the engine needs hazard_B = omega * hazard_A: shape matrices differ")
    ratio = (lA / lB) ** kA
    omega = float(np.mean(ratio))
    if not np.allclose(ratio, omega, rtol=1e-9, atol=0):
        raise ValueError("the engine needs hazard_B = omega * hazard_A with one omega for all entries")
    given = getattr(hazard_B, "omega", None)
    if given is not None and not np.isclose(float(given), omega, rtol=1e-9):
        raise ValueError("hazard_B.omega = %r disagrees with its scales (omega = %r)" % (given, omega))
    if not omega > 1.0:
        raise ValueError("omega must exceed 1")
    return omega


def emission_arrays(smjp_e):
    if hasattr(smjp_e, "matrix"):
        return smjp_e.matrix()
    K = len(smjp_e.state_space)
    E = np.ones((K, K))
    E[np.arange(K), np.arange(K)] = getattr(smjp_e, "how_likely_is_current_state_compared_to_others", 10)
    return E / E.sum(axis=1, keepdims=True)


def observation_arrays(data, state_space):
    """(times, symbol indices) sorted by time"""
    t = np.asarray(data.time, dtype=np.float64).reshape(-1)
    y = state_index(state_space, list(np.asarray(data.data).reshape(-1))) if len(t) else np.zeros(0, np.int32)
    order = np.argsort(t, kind="stable")
    return t[order], y[order]


def load_model(engine, state_space, hazard_A, hazard_B, smjp_e, t_end, obs_combine="sum"):
    omega = infer_omega(hazard_A, hazard_B)
    fg = getattr(smjp_e, "final_grid_time", t_end)
    if not np.isclose(fg, t_end):
        raise ValueError("emission.final_grid_time (%r) and t_end (%r) differ" % (fg, t_end))
    engine.set_model(hazard_A.shape_mat, hazard_A.scale_mat, omega, emission_arrays(smjp_e),
                     getattr(smjp_e, "likelihood_power", 1.0), float(t_end), obs_combine=obs_combine)
    return omega


# ------------------------------------------------------------------------------------------------
# the three sweep kernels under their reference names
# ------------------------------------------------------------------------------------------------
def sample_smjp_event_times(poisson_process, V, T, time_length, *, seed=None, sweep=0, mode="reference", engine=None):
    """W = sorted(thinned events U T) (smjp_utils.py:194-210).  `poisson_process` is the A_hat process:
    its mean (tau/lam_hat)^k is used as is.  mode='reference' places events like the reference
    (unit-scale Weibull density truncated to the sojourn), 'exact' proportionally to the hazard."""
    eng = engine or runtime.get_engine()
    states = list(poisson_process.state_space)
    K = len(states)
    shape = np.asarray(poisson_process.mean_function_params["shape"], dtype=np.float64)
    scale_hat = np.asarray(poisson_process.mean_function_params["scale"], dtype=np.float64)
    # candidates kernel draws N ~ Poisson((omega-1) (tau/scale)^k): omega = 2 with scale = lam_hat
    eng.set_model(shape, scale_hat, 2.0, np.full((K, K), 1.0 / K), 1.0, float(time_length))
    Vi = state_index(states, list(V))
    W = candidates_host(eng, [Vi], [np.asarray(T, dtype=np.float64)], runtime.next_seed() if seed is None else seed,
                        sweep=sweep, mode=mode)[0]
    return [float(w) for w in W]


def sample_smjp_trajectory_prior(hazard_A, pi_0, state_space, t_end, t_start=0, v_0=None, *, seed=None,
                                 honour_scale=False, engine=None):
    """competing-risk forward simulation (smjp_utils.py:216-254); returns (v list of labels, w list)."""
    from .chains import DeviceChains
    eng = engine or runtime.get_engine()
    states = list(state_space)
    K = len(states)
    if v_0 is None:
        pi0 = np.asarray(pi_0.prob_vector, dtype=np.float64)
    else:
        pi0 = np.zeros(K)
        pi0[state_index(states, [v_0])[0]] = 1.0
    span = float(t_end) - float(t_start)
    if not span > 0:
        raise ValueError("t_end must exceed t_start")
    eng.set_model(hazard_A.shape_mat, hazard_A.scale_mat, 2.0, np.full((K, K), 1.0 / K), 1.0, span)
    ch = DeviceChains(eng, 1, np.zeros(0)).init_prior(pi0, runtime.next_seed() if seed is None else seed,
                                                      honour_scale=honour_scale)
    V, T = ch.paths_host()
    w = (T[0] + float(t_start)).tolist()
    w[-1] = t_end
    return [states[i] for i in V[0]], w


def smjp_ffbs(W, data, state_space, hazard_A, hazard_B, smjp_e, t_end, u_str, **kw):
    return sample_smjp_trajectory_posterior(W, data, state_space, hazard_A, hazard_B, smjp_e, t_end, u_str, **kw)


def sample_smjp_trajectory_posterior(W, data, state_space, hazard_A, hazard_B, smjp_e, t_end, u_str, *,
                                     uniforms=None, seed=None, precision="f64", engine=None, full_output=False,
                                     obs_combine="sum"):
    """forward filter over the augmented states + backward sample + compaction (smjp_utils.py:261-330).
    Returns (s_states float64 labels, s_times, prob) with prob == 1.0 like the reference;
    full_output=True returns the result dict (messages, logZ, samples, ...) as a fourth item.
    obs_combine='sum' mirrors the reference (several observations in one grid interval are summed,
    SURVEY Q3); 'product' is the proper likelihood."""
    eng = engine or runtime.get_engine()
    states = list(state_space)
    load_model(eng, states, hazard_A, hazard_B, smjp_e, t_end, obs_combine)
    obs_t, obs_y = observation_arrays(data, states)
    Wa = np.asarray(W, dtype=np.float64)
    if uniforms is not None:
        uniforms = [np.asarray(uniforms, dtype=np.float64)]
    r = eng.ffbs_host([Wa], [obs_t], [obs_y], uniforms=uniforms, seed=runtime.next_seed() if seed is None else seed,
                      precision=precision, want_messages=full_output)
    raise_for_status(r["status"])
    smjp_e.augmented_state_space = enumerate_state_space(list(Wa), states)  # side effect of the reference (:278)
    labels = np.asarray([float(states[i]) if _is_number(states[i]) else states[i] for i in r["V"][0]])
    if labels.dtype != object:
        labels = labels.astype(np.float64)
    out = (labels, r["T"][0], 1.0)
    if full_output:
        n, K = len(Wa) - 1, len(states)
        r["alpha"] = triangular_to_dense(r["messages"][0], n, K)
        return out + (r,)
    return out


def _is_number(x):
    try:
        float(x)
        return True
    except (TypeError, ValueError):
        return False


def sample_data_posterior(V, T, state_space, emission, obs_times, *, seed=None, engine=None):
    """observations y ~ emission(v(t)) at obs_times (smjp_utils.py:350-364) -> sMJPDataWrapper"""
    from .chains import DeviceChains
    eng = engine or runtime.get_engine()
    states = list(state_space)
    K = len(states)
    obs_times = np.asarray(obs_times, dtype=np.float64)
    eng.set_model(np.ones((K, K)), np.ones((K, K)), 2.0, emission_arrays(emission), 1.0, float(T[-1]))
    ch = DeviceChains(eng, 1, obs_times).set_paths([state_index(states, list(V))], [np.asarray(T, dtype=np.float64)])
    ch.simulate_observations(runtime.next_seed() if seed is None else seed)
    y = ch.obs_y.cpu().numpy()
    inside = (obs_times >= T[0]) & (obs_times < T[-1])
    return sMJPDataWrapper(data=[states[i] for i in y[inside]], time=list(obs_times[inside]))
