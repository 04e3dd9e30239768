"""Evaluation helpers on the hot path's outputs (reference pkg/mcmc_utils.py): per-sample chain
metrics (:40-56), summaries (:58-92), the two-sample KS acceptance rule (:377-415) and the pickle
aggregate format (:436-451).  Plotting is out of scope."""
import pickle

import numpy as np
import scipy.stats as sss


def compute_evaluation_chain_metrics(agg, states):
    """state_times[s][i] = time sample i spends in s; state_jumps[s][i] = #entries of V equal to s
    (the first entry and the duplicated end entry included)."""
    V_list, T_list = agg["V"], agg["T"]
    n = len(T_list)
    state_times = {s: [0 for _ in range(n)] for s in states}
    state_jumps = {s: [0 for _ in range(n)] for s in states}
    for i, (V, T) in enumerate(zip(V_list, T_list)):
        V = np.asarray(V)
        dt = np.diff(np.asarray(T, dtype=np.float64))
        for s in states:
            hit = np.isclose(V.astype(np.float64), float(s)) if V.dtype != object else (V == s)
            state_jumps[s][i] = int(np.sum(hit))
            state_times[s][i] += float(np.sum(dt[hit[:-1]]))
    return state_times, state_jumps


def metrics_from_arrays(time_in_state, jumps, states):
    """batched engine output [C,K] -> the dict layout of compute_evaluation_chain_metrics"""
    return ({s: list(np.asarray(time_in_state)[:, k]) for k, s in enumerate(states)},
            {s: list(np.asarray(jumps)[:, k]) for k, s in enumerate(states)})


def compute_metric_summaries(samples, states):
    return {name: {s: float(fn(samples[s])) for s in states}
            for name, fn in (("median", np.median), ("mean", np.mean), ("var", np.var))}


def compute_metric_summaries_trajectory(state_times, state_jumps, states):
    return compute_metric_summaries(state_times, states), compute_metric_summaries(state_jumps, states)


def compute_ks_twosample_time_jump(time_info, time_info_prior, jump_info, jump_info_prior, state_space, skip=30,
                                   verbose=True):
    """two-sample KS per state on time-in-state and jump counts, thinned by `skip`; reject when
    D > 1.224 sqrt((n+m)/(nm)) (alpha = 0.05).  Returns dict(ub, times_ks, jumps_ks, reject_times,
    reject_jumps) in addition to printing what the reference prints."""
    n = len(time_info[state_space[0]][::skip])
    m = len(time_info_prior[state_space[0]][::skip])
    ub = 1.224 * np.sqrt((n + m) / (n * m))
    K = len(state_space)
    times_ks, jumps_ks = np.zeros((K, 2)), np.zeros((K, 2))
    for k, s in enumerate(state_space):
        r = sss.ks_2samp(time_info[s][::skip], time_info_prior[s][::skip])
        times_ks[k] = r.statistic, r.pvalue
        r = sss.ks_2samp(jump_info[s][::skip], jump_info_prior[s][::skip])
        jumps_ks[k] = r.statistic, r.pvalue
    out = {"ub": ub, "times_ks": times_ks, "jumps_ks": jumps_ks, "reject_times": times_ks[:, 0] > ub,
           "reject_jumps": jumps_ks[:, 0] > ub}
    if verbose:
        print("[skip: {}]".format(skip))
        print(" ---> KS Results <--- ")
        print("Reject if D_{{n,m}} > {}".format(ub))
        print("--> times_ks\n[stat,pvalue,do_we_reject?]")
        print(np.c_[times_ks, out["reject_times"]])
        print("--> jumps_ks\n[stat,pvalue,do_we_reject?]")
        print(np.c_[jumps_ks, out["reject_jumps"]])
    return out


def save_samples_in_pickle(aggregate, omega, expName, uuid_str, n_iters=None):
    fn = "results_{}_{}_{}.pkl".format(expName, uuid_str, "final" if n_iters is None else n_iters)
    with open(fn, "wb") as f:
        pickle.dump({"agg": aggregate, "uuid_str": uuid_str, "omega": omega}, f)
    return fn


def load_samples_in_pickle(fn):
    with open(fn, "rb") as f:
        d = pickle.load(f)
    return d["agg"], d["uuid_str"], d["omega"]


# ------------------------------------------------------------------------------------------------
# Columnar store for many chain-sweeps.  The reference pickles Python lists of
# ragged arrays (mcmc_utils.py:436-451), which cannot hold a 65 536-chain run; here every ragged list
# becomes one flat array plus int64 offsets inside a single .npz, and scalars / dense arrays are
# stored as they are.  load_aggregate_columnar returns the same dict-of-lists the pickle held.
# ------------------------------------------------------------------------------------------------
def save_aggregate_columnar(fn, aggregate, omega=None, uuid_str=None):
    """aggregate: dict whose values are lists of 1-D arrays (ragged: 'W', 'T', 'V', ...), lists of scalars
    ('prob'), dense arrays (e.g. time_in_state [S,C,K]) or nested dicts of the same (e.g. 'prior')."""
    cols = {}

    def put(prefix, value):
        if isinstance(value, dict):
            cols[prefix + "@dict"] = np.array(sorted(value.keys()))
            for k, v in value.items():
                put(prefix + "/" + k, v)
        elif isinstance(value, (list, tuple)) and len(value) and all(np.ndim(x) == 1 for x in value):
            lens = np.array([len(x) for x in value], dtype=np.int64)
            cols[prefix + "@offs"] = np.concatenate([[0], np.cumsum(lens)])
            cols[prefix + "@flat"] = np.concatenate([np.asarray(x) for x in value]) if lens.sum() else np.zeros(0)
        else:
            cols[prefix + "@dense"] = np.asarray(value)

    put("agg", aggregate)
    if omega is not None:
        cols["omega"] = np.float64(omega)
    if uuid_str is not None:
        cols["uuid_str"] = np.array(str(uuid_str))
    np.savez_compressed(fn, **cols)
    return fn if str(fn).endswith(".npz") else str(fn) + ".npz"


def load_aggregate_columnar(fn):
    """inverse of save_aggregate_columnar: returns (aggregate, uuid_str, omega) like load_samples_in_pickle"""
    z = np.load(fn, allow_pickle=False)

    def get(prefix):
        if prefix + "@dict" in z:
            return {str(k): get(prefix + "/" + str(k)) for k in z[prefix + "@dict"]}
        if prefix + "@offs" in z:
            offs, flat = z[prefix + "@offs"], z[prefix + "@flat"]
            return [flat[offs[i]: offs[i + 1]] for i in range(len(offs) - 1)]
        d = z[prefix + "@dense"]
        return d.tolist() if d.ndim == 1 and d.dtype.kind == "f" and prefix.endswith("prob") else d

    agg = get("agg")
    return agg, (str(z["uuid_str"]) if "uuid_str" in z else None), (float(z["omega"]) if "omega" in z else None)
