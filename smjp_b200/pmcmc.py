"""Drop-in for the reference's pkg/pmcmc.py: particle-marginal Metropolis-Hastings around the
bootstrap particle filter (pmcmc.py:8-74, smjp_smc :224-310).

smjp_smc runs on the GPU (smjp_pf_host).  Two reference defects are NOT mirrored:
resampled particles share Python lists there, and the per-particle clock leaks between particles."""
import numpy as np

from . import runtime
from .engine import pf_host, raise_for_status
from .mcmc_utils import load_samples_in_pickle, save_samples_in_pickle
from .smjp_utils import emission_arrays, observation_arrays
from .utils import np_log


def smjp_smc(*args, **kwargs):
    """keywords as the reference: pi_0, time_final, N, data, A, emission, states, uuid; extras:
    seed, resampler ('multinomial' = reference, 'systematic'), family ('weibull' = reference, 'gamma': shape_mat
    holds the gamma shapes and scale_mat the gamma scales), extend, engine.
    Returns (path_states, path_times, log_likelihood)."""
    states = list(kwargs["states"])
    hazard_A, emission = kwargs["A"], kwargs["emission"]
    eng = kwargs.get("engine") or runtime.get_engine()
    K = len(states)
    eng.set_model(hazard_A.shape_mat, hazard_A.scale_mat, 2.0, emission_arrays(emission), 1.0, float(kwargs["time_final"]))
    obs_t, obs_y = observation_arrays(kwargs["data"], states)
    seed = kwargs.get("seed")
    r = pf_host(eng, obs_t, obs_y, np.asarray(kwargs["pi_0"].prob_vector, dtype=np.float64), int(kwargs["N"]),
                runtime.next_seed() if seed is None else seed, resampler=kwargs.get("resampler", "multinomial"), C=1,
                family=kwargs.get("family", "weibull"), extend=kwargs.get("extend", False))
    raise_for_status(r["status"])
    return [states[i] for i in r["V"][0]], list(r["T"][0]), float(r["loglik"][0])


def parameter_proposal(states):
    K = len(states)
    return {"shape": runtime.rng().uniform(0.6, 3, K * K).reshape(K, K)}, 0


def update_parameters(smc_input, params):
    smc_input["A"].update_parameters(params)


def pmcmc(inference, number_of_particles, number_of_samples, save_iter, states,
          hazard_A, emission, time_final,
          data, pi_0, uuid_str, omega, filename, load_file, *, seed=None, resampler="multinomial", save=True):
    """Returns (aggregate {'V','T','theta','prob'}, uuid_str) (pmcmc.py:8-74)."""
    if load_file:
        aggregate, uuid_str, _ = load_samples_in_pickle(filename)
        return aggregate, uuid_str
    if seed is not None:
        runtime.seed(seed)
    smc_input = {"pi_0": pi_0, "time_final": time_final, "N": number_of_particles, "data": data, "A": hazard_A,
                 "emission": emission, "states": states, "uuid": uuid_str, "resampler": resampler}
    params = {"shape": hazard_A.shape_mat}
    V, T, ldata = smjp_smc(**smc_input)
    lprop = 0
    aggregate = {"V": [], "T": [], "theta": [], "prob": []}
    for i in range(number_of_samples):
        params_prop, lprop_prop = parameter_proposal(states)
        update_parameters(smc_input, params_prop)
        V_prop, T_prop, ldata_prop = smjp_smc(**smc_input)
        coin_flip = np_log(runtime.rng().uniform(0, 1))[0]
        alpha = ldata_prop + lprop - ldata - lprop_prop  # pmcmc.py:59
        if coin_flip <= alpha:
            V, T, params, ldata, lprop = V_prop, T_prop, params_prop, ldata_prop, lprop_prop
        if save and i % 300 == 0:
            save_samples_in_pickle(aggregate, None, "pmcmc", uuid_str, i)
        aggregate["V"].append(V)
        aggregate["T"].append(T)
        aggregate["theta"].append(params)
        aggregate["prob"].append(ldata)
    if save:
        save_samples_in_pickle(aggregate, None, "pmcmc", uuid_str)
    return aggregate, uuid_str
