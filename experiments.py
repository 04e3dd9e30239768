"""Driver mirroring the reference's experiments.py::experiment_2 (the default run of main.py): the
3-state Weibull sMJP of experiments.py:154-222, sampled with Rao-Teh and with particle MCMC on the
B200 engine, compared by the reference's KS report.  Plots are out of scope (headless)."""
import uuid

import numpy as np

from smjp_b200.distributions import MultinomialDistribution
from smjp_b200.mcmc_utils import compute_evaluation_chain_metrics, compute_ks_twosample_time_jump, compute_metric_summaries_trajectory
from smjp_b200.pmcmc import pmcmc
from smjp_b200.raoteh import raoteh
from smjp_b200.smjp_utils import (PoissonProcess, create_upperbound_scale, sMJPDataWrapper, smjpEmission,
                                  smjpHazardFunction)


def build_experiment_2_model(likelihood_power=1.0):
    """model construction of experiments.py:154-222, verbatim in values"""
    state_space = [1, 2, 3]
    time_final = 2.0
    omega = 2
    obs_times = np.arange(0.1, time_final, 0.1)
    shape_mat = np.array([[1.606, 1.933, 0.865],
                          [1.938, 0.869, 1.751],
                          [1.69, 0.64, 0.696]])
    scale_mat = np.ones((3, 3))
    scale_mat_tilde = create_upperbound_scale(shape_mat, scale_mat, omega)
    scale_mat_hat = create_upperbound_scale(shape_mat, scale_mat, omega - 1)
    hazard_A = smjpHazardFunction(state_space, shape_mat, scale_mat)
    hazard_B = smjpHazardFunction(state_space, shape_mat, scale_mat_tilde, omega=omega)
    hazard_A_hat = smjpHazardFunction(state_space, shape_mat, scale_mat_hat)
    poisson_process_A_hat = PoissonProcess(state_space, None, hazard_A_hat, {"shape": shape_mat, "scale": scale_mat_hat})
    poisson_process_B = PoissonProcess(state_space, None, hazard_B, {"shape": shape_mat, "scale": scale_mat_tilde})
    emission = smjpEmission(state_space, poisson_process_B, time_final, likelihood_power)
    data_samples = np.array([1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 3, 3, 2, 1, 1, 3])
    data = sMJPDataWrapper(data=data_samples, time=obs_times)
    pi_0 = MultinomialDistribution({"prob_vector": np.ones(3) / 3, "translation": state_space})
    return dict(state_space=state_space, time_final=time_final, omega=omega, obs_times=obs_times, hazard_A=hazard_A,
                hazard_B=hazard_B, poisson_process_A_hat=poisson_process_A_hat, poisson_process_B=poisson_process_B,
                emission=emission, data=data, pi_0=pi_0)


def experiment_2(likelihood_power=1., inference=("trajectory",), number_of_samples=6000, number_of_particles=10,
                 save_iter=300, load_file=False, rt_filename=None, pm_filename=None, seed=None, save=False, skip=30):
    """Rao-Teh vs particle MCMC on the default model; returns the KS report dict."""
    inference = list(inference)
    m = build_experiment_2_model(likelihood_power)
    uuid_str = uuid.uuid4()
    rt_agg, rt_uuid, _ = raoteh(inference, number_of_samples, save_iter, m["state_space"], m["time_final"], m["emission"],
                                m["data"], m["pi_0"], m["hazard_A"], m["hazard_B"], m["poisson_process_A_hat"],
                                m["poisson_process_B"], uuid_str, m["omega"], m["obs_times"], rt_filename, load_file,
                                seed=seed, save=save)
    m2 = build_experiment_2_model(likelihood_power)  # pmcmc mutates hazard_A: give it its own objects
    pm_agg, pm_uuid = pmcmc(inference, number_of_particles, number_of_samples, save_iter, m2["state_space"], m2["hazard_A"],
                            m2["emission"], m2["time_final"], m2["data"], m2["pi_0"], uuid_str, m2["omega"], pm_filename,
                            load_file, seed=None if seed is None else seed + 1, save=save)
    rt_t, rt_j = compute_evaluation_chain_metrics(rt_agg, m["state_space"])
    pm_t, pm_j = compute_evaluation_chain_metrics(pm_agg, m["state_space"])
    print("raoteh :", compute_metric_summaries_trajectory(rt_t, rt_j, m["state_space"])[0]["mean"])
    print("pmcmc  :", compute_metric_summaries_trajectory(pm_t, pm_j, m["state_space"])[0]["mean"])
    return compute_ks_twosample_time_jump(rt_t, pm_t, rt_j, pm_j, m["state_space"], skip=skip)
