"""Drop-in for the reference's pkg/raoteh.py: the Rao-Teh Gibbs loop (raoteh.py:7-107), plus the
batched, device-resident form the reference lacks (raoteh_batch).

Each iteration is   W <- sample_smjp_event_times(A_hat, V, T);   (V, T) <- smjp_ffbs(W, data)
(raoteh.py:58-59); with 'parameters' in `inference` the shapes are then redrawn from U(0.6, 3) and
pushed into hazard_A / hazard_B / both Poisson processes exactly as the reference does (:74-78,
:109-115, :173-193 -- its grid-search MLE is computed and discarded there; here the
placeholder draw alone is kept and hold_times.txt is not written).
"""
import numpy as np

from . import runtime
from .mcmc_utils import load_samples_in_pickle, save_samples_in_pickle
from .smjp_utils import (create_upperbound_scale, sample_data_posterior, sample_smjp_event_times,
                         sample_smjp_trajectory_prior, smjp_ffbs)


def inference_error_checking(inference):
    if len(inference) == 0:
        raise ValueError("We need to infer something.")
    if "trajectory" not in inference:
        raise ValueError("We always need to sample trajectory.")


def sample_smjp_parameters(data, V, T, state_space, sigma=1.0):
    """the reference's parameter step: a U(0.6, 3) draw per (v, v') (raoteh.py:115)"""
    K = len(state_space)
    return {"shape": runtime.rng().uniform(0.6, 3.0, K * K).reshape(K, K)}


def filter_T_by_state(T, V, curr_s, next_s):
    """holding times of the consecutive path entries (curr_s, next_s) (raoteh.py:136-151), as a list"""
    V = np.asarray(V)
    T = np.asarray(T, dtype=np.float64)
    sel = np.nonzero((V[:-1] == curr_s) & (V[1:] == next_s))[0]
    return list(T[sel + 1] - T[sel])


def esimtate_weibull_shape(T, grid=0.001, grid_max=5):
    """(k-hat, g(k-hat)) of the reference's grid search (raoteh.py:153-167) -- on the device: the holding
    times are laid out as one single-state path, whose only transition (0, 0) carries them"""
    from .engine import shape_mle_host
    h = np.asarray(T, dtype=np.float64)
    if len(h) == 0:
        raise ValueError("no holding times")
    r = shape_mle_host(runtime.get_engine(), [np.zeros(len(h) + 1, np.int32)], [np.concatenate([[0.0], np.cumsum(h)])],
                       1, True, grid, grid_max)
    return float(r["khat"][0, 0]), float(r["g"][0, 0])


def esimtate_weibull_shape_mat(V, T, state_space):
    """K x K shape estimates from the holding times of one trajectory (raoteh.py:119-134): one kernel launch
    for all (v, v'); transitions that never occur get the reference's U(0.6, 3) draw (hold_times.txt is not
    written)"""
    from .engine import shape_mle_host
    labels = list(state_space)
    idx = np.asarray([labels.index(x) for x in np.asarray(V).tolist()], dtype=np.int32)
    r = shape_mle_host(runtime.get_engine(), [idx], [np.asarray(T, dtype=np.float64)], len(labels), True)
    shapes = r["khat"].copy()
    empty = r["count"] == 0
    shapes[empty] = runtime.rng().uniform(0.6, 3.0, int(empty.sum()))
    return shapes


def update_parameters(hazard_A, hazard_B, poisson_proc_A_hat, poisson_proc_B, theta):
    """push new shapes into A, B = omega*A, A_hat = (omega-1)*A and the B process (raoteh.py:173-193)"""
    omega = hazard_B.omega
    hazard_A.update_parameters(theta)
    tilde = create_upperbound_scale(hazard_A.shape_mat, hazard_A.scale_mat, omega)
    hat = create_upperbound_scale(hazard_A.shape_mat, hazard_A.scale_mat, omega - 1)
    hazard_B.update_parameters({"shape": hazard_A.shape_mat, "scale": tilde})
    poisson_proc_A_hat.update_parameters({"shape": hazard_A.shape_mat, "scale": hat})
    if getattr(poisson_proc_A_hat, "hazard_function", None) is not None:
        poisson_proc_A_hat.hazard_function.update_parameters({"shape": hazard_A.shape_mat, "scale": hat})
    poisson_proc_B.update_parameters({"shape": hazard_A.shape_mat, "scale": tilde})


def _is_number(x):
    try:
        float(x)
        return True
    except (TypeError, ValueError):
        return False


def _a_hat_is_scaled_hazard(hazard_A, hazard_B, poisson_process_A_hat):
    """True when the candidate process is (omega - 1) * hazard_A (experiments.py:177-205): same shapes, scale
    lambda / (omega - 1)^(1/k) -- the engine's sweep draws candidates from exactly that process"""
    try:
        omega = float(hazard_B.omega)
        kA = np.asarray(hazard_A.shape_mat, dtype=np.float64)
        kP = np.asarray(poisson_process_A_hat.mean_function_params["shape"], dtype=np.float64)
        lP = np.asarray(poisson_process_A_hat.mean_function_params["scale"], dtype=np.float64)
        want = create_upperbound_scale(kA, hazard_A.scale_mat, omega - 1.0)
        return kA.shape == kP.shape and np.allclose(kA, kP, rtol=1e-12, atol=0) and np.allclose(lP, want, rtol=1e-9, atol=0)
    except Exception:
        return False


def raoteh(inference, number_of_samples, save_iter, state_space, time_final, emission,
           data, pi_0, hazard_A, hazard_B, poisson_process_A_hat, poisson_process_B,
           uuid_str, omega, obs_times, filename, load_file, *, seed=None, verbose=False, save=True,
           candidate_mode="reference", precision="f64", obs_combine="sum", fused_sweep=True):
    """Returns (aggregate {'W','T','V','prob'[,'theta'][,'prior']}, uuid_str, omega)."""
    inference_error_checking(inference)
    if load_file:
        return load_samples_in_pickle(filename)
    if seed is not None:
        runtime.seed(seed)
    V, T = sample_smjp_trajectory_prior(hazard_A, pi_0, state_space, time_final)
    aggregate = {"W": [], "T": [], "V": [], "prob": []}
    if "parameters" in inference:
        aggregate["theta"] = []
    if "prior" in inference:
        aggregate["prior"] = {"W": [], "T": [], "V": []}
    if "data" in inference:
        aggregate["data"] = []
    # One fused host call per iteration (smjp_sweep_host: candidates -> FFBS in one launch sequence, model uploaded
    # once) whenever the A_hat process IS (omega - 1) * hazard_A, as experiments.py:177-205 builds it; otherwise -- or
    # with 'data' inference, which changes the observations every iteration -- the two reference calls are made.
    from .engine import raise_for_status, sweep_host
    from .smjp_utils import (enumerate_state_space, load_model, observation_arrays, state_index)
    eng = runtime.get_engine()
    states = list(state_space)
    fused = fused_sweep and "data" not in inference and _a_hat_is_scaled_hazard(hazard_A, hazard_B, poisson_process_A_hat)
    if fused:
        load_model(eng, states, hazard_A, hazard_B, emission, time_final, obs_combine)
        obs_t, obs_y = observation_arrays(data, states)
        Vi, Ta = state_index(states, list(V)), np.asarray(T, dtype=np.float64)
        labels = np.asarray([float(x) for x in states]) if all(_is_number(x) for x in states) else np.asarray(states, dtype=object)
    for i in range(number_of_samples):
        if "data" in inference:
            data = sample_data_posterior(V, T, state_space, emission, obs_times)
            aggregate["data"].append(data)
        if fused:
            r = sweep_host(eng, [Vi], [Ta], obs_t, obs_y, runtime.next_seed(), sweep=i, mode=candidate_mode, precision=precision)
            raise_for_status(r["status"])
            Wa = r["W"][0]
            W = Wa.tolist()
            Vi, Ta = r["V"][0], r["T"][0]
            V = labels[Vi]  # float64 labels for numeric state spaces (smjp_utils.py:328)
            T, prob = Ta, 1.0
            # side effect of the reference (:278): augmented states = states x grid[:-1], state-major
            emission.augmented_state_space = np.column_stack([np.repeat(labels, len(Wa) - 1), np.tile(Wa[:-1], len(states))])
        else:
            W = sample_smjp_event_times(poisson_process_A_hat, V, T, time_final, mode=candidate_mode)
            V, T, prob = smjp_ffbs(W, data, state_space, hazard_A, hazard_B, emission, time_final,
                                   "raoteh_{}_{}".format(uuid_str, i), precision=precision, obs_combine=obs_combine)
        aggregate["W"].append(W)
        aggregate["V"].append(V)
        aggregate["T"].append(T)
        aggregate["prob"].append(prob)
        if "parameters" in inference:
            theta = sample_smjp_parameters(data, V, T, state_space)
            aggregate["theta"].append(theta)
            update_parameters(hazard_A, hazard_B, poisson_process_A_hat, poisson_process_B, theta)
            if fused:  # the hazards changed in place: upload them (raoteh.py:74-78 re-reads them every sweep)
                load_model(eng, states, hazard_A, hazard_B, emission, time_final, obs_combine)
        if "prior" in inference:
            _V, _T = sample_smjp_trajectory_prior(hazard_A, pi_0, state_space, time_final)
            aggregate["prior"]["V"].append(_V)
            aggregate["prior"]["T"].append(_T)
            if fused:  # the prior sampler loaded its own model into the engine
                load_model(eng, states, hazard_A, hazard_B, emission, time_final, obs_combine)
        if verbose:
            print("i = {}  |W| = {}  |T| = {}".format(i, len(W), len(T)))
        if save and (i % save_iter) == 0 and i > 0:
            save_samples_in_pickle(aggregate, omega, "raoteh", uuid_str, i)
    if save:
        save_samples_in_pickle(aggregate, omega, "raoteh", uuid_str)
    return aggregate, uuid_str, omega


def raoteh_batch(number_of_chains, number_of_sweeps, state_space, time_final, emission, data, pi_0, hazard_A,
                 hazard_B, *, seed=0, burn_in=0, candidate_mode="reference", precision="f64", engine=None,
                 keep_paths=False, chain_base=0, obs_combine="sum"):
    """`number_of_chains` independent Rao-Teh chains advanced together on one GPU; the paths never
    leave the device.  Returns dict(time_in_state [S,C,K], jumps [S,C,K] for the S kept sweeps,
    paths (last sweep, or every kept sweep when keep_paths), state_steps)."""
    from .chains import DeviceChains
    from .smjp_utils import load_model, observation_arrays
    eng = engine or runtime.get_engine()
    states = list(state_space)
    load_model(eng, states, hazard_A, hazard_B, emission, time_final, obs_combine)
    obs_t, obs_y = observation_arrays(data, states)
    ch = DeviceChains(eng, number_of_chains, obs_t, chain_base=chain_base)
    ch.init_prior(np.asarray(pi_0.prob_vector, dtype=np.float64), seed).set_observations(obs_y)
    tis, jm, paths, steps = [], [], [], 0.0
    for sw in range(burn_in + number_of_sweeps):
        ch.sweep(seed=seed, sweep=sw, mode=candidate_mode, precision=precision)
        if sw >= burn_in:
            tis.append(ch.time_in_state.cpu().numpy().copy())
            jm.append(ch.jumps.cpu().numpy().copy())
            steps += ch.state_steps_last_sweep()[0]
            if keep_paths:
                paths.append(ch.paths_host())
    bad = int((ch.status != 0).sum().item())
    if bad:
        from .engine import raise_for_status
        raise_for_status(ch.status.cpu().numpy())
    return {"time_in_state": np.array(tis), "jumps": np.array(jm), "paths": paths if keep_paths else ch.paths_host(),
            "state_steps": steps, "chains": ch}
