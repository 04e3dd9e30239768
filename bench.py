#!/usr/bin/env python
"""Benchmark of the Rao-Teh Gibbs sweep (candidates -> forward filter -> backward sample) on B200.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): 4096 independent chains PER GPU of the 3-state Weibull sMJP of
experiments.py:172-174, T=100, Omega=2, noisy discrete observations every 0.1 (D=999) simulated from a
prior path per chain; chains start from an independent prior draw.  A "step" is one Gibbs sweep of
every chain.  Metric: live augmented state-steps/s (K*n(n+1)/2 per chain-sweep),
whole-job aggregate over all GPUs; chain-sweeps/s and grid-steps/s are reported beside it.
One JSON line on stdout (rank 0).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = np.array([[1.606, 1.933, 0.865], [1.938, 0.869, 1.751], [1.69, 0.64, 0.696]])  # experiments.py:172-174
K = 3
OMEGA = 2.0
T_END = 100.0
OBS_DT = 0.1
CHAINS_PER_GPU = 4096
SEED = 20261019
SWEEPS_PER_REF_STEP = 4
METRIC = "ffbs_state_steps_per_sec"
UNIT = "state-steps/s"


def workload_name(chains):
    return "cfg2: %d chains/GPU, K=3 Weibull sMJP (experiments.py:172-174), T=100, Omega=2, D=999 obs" % chains


def algorithmic_bytes(n_per_chain, D, esz, k=K):
    """per chain-sweep 2*sizeof(real) per live state-step (message written once by the
    forward pass, read once by the backward pass) + 40 B per grid step (W, uniform, index, logZ, path)
    + 12 B per observation.  NOMINAL count: every state (v, W[j]), j <= i, of every row i."""
    n = n_per_chain.astype(np.float64)
    return float(np.sum(2 * esz * k * n * (n + 1) / 2 + 40 * n + 12 * D))


def retained_bytes(live_pairs, n_per_chain, D, esz, k=K):
    """SURVEY.md 8(d) with a live window: "the count is the states actually retained" -- the same formula with the
    K * (grid pairs the kernels processed) in place of K n(n+1)/2 (device counter ffbs_live_pairs)."""
    n = n_per_chain.astype(np.float64)
    return float(2 * esz * k * live_pairs + np.sum(40 * n + 12 * D))


def roofline_block(kernel, retained, nominal, launches, kernel_ms, peak, peak_src, live_fraction, traffic, **extra):
    """`achieved` / `frac`: RETAINED algorithmic bytes per launch / the kernel's average launch duration (CUDA events
    on the launching stream) / measured HBM peak.  The nominal K n(n+1)/2 figure (what an engine without the live
    window would have to move) is kept beside it under `nominal_*`; it is not a bandwidth."""
    launches = max(launches, 1)
    k_s = kernel_ms / launches / 1e3
    ach = retained / launches / k_s / 1e9
    nom = nominal / launches / k_s / 1e9
    r = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
         "peak_source": peak_src, "kernel": kernel, "kernel_ms_avg": kernel_ms / launches,
         "algorithmic_bytes_per_launch": retained / launches, "live_fraction": live_fraction,
         "nominal_bytes_per_launch": nominal / launches, "nominal_achieved": nom, "nominal_frac": nom / peak,
         "traffic_over_algorithmic": (traffic / (retained / launches)) if traffic else None}
    r.update(extra)
    return r


# ------------------------------------------------------------------------------------------------
SAMPLER_SRC = r"""
import json, sys, time, select
idx = int(sys.argv[1])
period = float(sys.argv[2])
out = {"sm": [], "smmax": None, "mask": 0, "via": "nvml"}
try:
    import pynvml as n
    n.nvmlInit()
    h = n.nvmlDeviceGetHandleByIndex(idx)
    out["smmax"] = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
    def reasons():
        try:
            return int(n.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            return int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
    n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM); reasons()      # first queries of each kind are slow
    print("ready", flush=True)
    sys.stdin.readline()                                         # "b": begin
    while True:
        out["sm"].append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
        out["mask"] |= reasons()
        r, _, _ = select.select([sys.stdin], [], [], period)
        if r:
            break
except Exception as e:
    out["error"] = repr(e)
    print("ready", flush=True)
print(json.dumps(out), flush=True)
"""


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region by a helper PROCESS with its own NVML
    session (started and primed before the region, sampling every ~80 ms between begin() and end(): every NVML
    query holds a driver lock for a few ms, whichever process issues it, so there are few of them).  NVML
    queries take 5-50 ms now and then: issued from this process -- inline or from a thread -- they were
    measured to stall the launch path of the timed loop by that much; a `nvidia-smi` subprocess per sample
    costs ~0.2 s.  The helper does not touch this process."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.res = None
        self.proc = None
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", SAMPLER_SRC, str(index), os.environ.get("SMJP_BENCH_SAMPLER_S", "0.08")], stdin=subprocess.PIPE,
                                         stdout=subprocess.PIPE, text=True)
            self.proc.stdout.readline()  # "ready"
        except Exception:
            self.proc = None

    def begin(self):
        if self.proc:
            try:
                self.proc.stdin.write("b\n")
                self.proc.stdin.flush()
            except Exception:
                pass

    def end(self):
        if self.proc and self.res is None:
            try:
                self.proc.stdin.write("e\n")
                self.proc.stdin.flush()
                self.res = json.loads(self.proc.stdout.readline())
                self.proc.wait(timeout=5)
            except Exception:
                self.res = {"sm": [], "smmax": None, "mask": 0, "error": "sampler process failed"}

    def summary(self):
        self.end()
        r = self.res or {"sm": [], "smmax": None, "mask": 0}
        if not r["sm"]:
            return {"sm_mhz": None, "sm_max_mhz": r.get("smmax"), "reasons": [], "samples": 0, "error": r.get("error")}
        return {"sm_mhz": float(np.median(r["sm"])), "sm_max_mhz": r.get("smmax"),
                "reasons": sorted(name for bit, name in self.REASONS.items() if r["mask"] & bit),
                "samples": len(r["sm"]), "via": "nvml (helper process)"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


FFBS_KERNEL_NAMES = {0: "ffbs_kernel", 1: "ffbs_item_kernel", 2: "ffbs_big_kernel", 3: "ffbs_warp_kernel",
                     4: "ffbs_fast_kernel", 5: "ffbs_fastk_kernel"}  # smjp_ctx_get_stat("ffbs_kernel")


def captured_traffic(traffic, key, kernel):
    """DRAM bytes per launch from the committed ncu capture -- only when it is a capture of the kernel that ran"""
    t = ((traffic or {}).get(key) or {})
    return t.get("dram_bytes_per_launch") if str(t.get("kernel", "")).startswith(kernel + "<") else None


def ceilings_from_capture(t):
    """The compute ceilings of a kernel from its committed ncu --set full capture (profiles/roofline_traffic.json):
    the launch cannot be shorter than its duration x the busiest pipe's utilisation; `ceiling_over_measured` is
    that fraction -- 1.0 would mean the kernel sits on its arithmetic / issue ceiling."""
    if not t or "issue_active_pct" not in t:
        return None
    busiest = max(float(t["issue_active_pct"]), float(t.get("pipe_active_pct_ncu", 0.0)))
    return {"pipe": t.get("pipe"), "pipe_active_pct_ncu": t.get("pipe_active_pct_ncu"),
            "issue_active_pct_ncu": t["issue_active_pct"], "thread_inst_per_pair": t.get("pipe_thread_inst_per_pair"),
            "warp_inst_per_pair_all_pipes": t.get("warp_inst_per_warp_pair"),
            "kernel_ms_at_ceiling_ncu": float(t["gpu_time_ms"]) * busiest / 100.0, "ceiling_over_measured": busiest / 100.0,
            "source": "profiles/roofline_traffic.json (ncu --set full)"}


def ncu_traffic():
    """per-launch DRAM bytes of the FFBS kernel from the committed ncu --set full capture, if any"""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return None


# ------------------------------------------------------------------------------------------------
# CPU legs (the ONLY place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_chain_sweeps(args):
    """one worker: run `sweeps` oracle sweeps on chain `c`; returns (state_steps, grid_steps, seconds)"""
    c, sweeps, t_end = args
    from oracle import philox, smjp_oracle as O
    emis = O.emission_matrix(K)
    scale = np.ones((K, K))
    obs_t = np.arange(OBS_DT, t_end, OBS_DT)
    V0, T0 = O.prior_path(SHAPE, scale, np.ones(K) / K, t_end, seed=SEED + 1, chain=c)
    y = O.simulate_observations(V0, T0, obs_t, emis, seed=SEED + 2, chain=c)
    V, T = O.prior_path(SHAPE, scale, np.ones(K) / K, t_end, seed=SEED + 3, chain=c)
    ss = gs = 0.0
    t0 = time.perf_counter()
    for sw in range(sweeps):
        W = O.candidates(V, T, SHAPE, scale, OMEGA, seed=SEED, sweep=sw, chain=c)
        n = len(W) - 1
        u = philox.u01(SEED, philox.TAG_BACKWARD, np.arange(n), 0, sw, c)
        r = O.ffbs(W, obs_t, y, SHAPE, scale, OMEGA, emis, 1.0, t_end, u)
        V, T = r["V"], r["T"]
        ss += K * n * (n + 1) / 2
        gs += n
    return ss, gs, time.perf_counter() - t0


def cpu_sample(n_chains, sweeps, t_end, procs):
    """oracle port on `procs` host processes, one chain each at a time"""
    import multiprocessing as mp
    jobs = [(c, sweeps, t_end) for c in range(n_chains)]
    t0 = time.perf_counter()
    if procs <= 1:
        res = [_cpu_chain_sweeps(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_cpu_chain_sweeps, jobs)
    wall = time.perf_counter() - t0
    ss = sum(r[0] for r in res)
    gs = sum(r[1] for r in res)
    return ss, gs, wall


def time_unmodified_reference(budget_s=12.0):
    """The UNMODIFIED reference (oracle/_ref = the bytecode `make -C oracle` compiles from /root/reference's
    pure-Python modules, imported sourceless behind oracle/ref_shim.py) on cfg1 = main.py's default model (experiments.py:154-222): the loop body of raoteh.py:58-59
    (candidates + smjp_ffbs), one core -- BASELINE.md section 3 item 1.  ~1 sweep/s, so a handful of sweeps."""
    try:
        from oracle import ref_shim
        if not ref_shim.reference_available():
            return {"unavailable": "oracle/_ref is missing (run `make -C oracle` where /root/reference exists)"}
        ref = ref_shim.load_reference()
        shape = SHAPE
        model = ref_shim.build_reference_model(ref, [1, 2, 3], shape, np.ones((3, 3)), 2, 2.0, np.arange(0.1, 2.0, 0.1),
                                               [1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 3, 3, 2, 1, 1, 3], 1.0)
        np.random.seed(SEED % (2 ** 31))
        with ref_shim.quiet():
            V, T = ref.smjp.sample_smjp_trajectory_prior(model.hazard_A, model.pi_0, [1, 2, 3], 2.0)
        ss = gs = 0.0
        sweeps = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < budget_s and sweeps < 200:
            with ref_shim.quiet():
                W = ref.smjp.sample_smjp_event_times(model.pp_A_hat, V, T, 2.0)
                V, T, _ = ref.smjp.smjp_ffbs(W, model.data, [1, 2, 3], model.hazard_A, model.hazard_B, model.emission, 2.0, "bench")
            n = len(W) - 1
            ss += 3 * n * (n + 1) / 2
            gs += n
            sweeps += 1
        dt = time.perf_counter() - t0
        return {"kind": "reference", "value": ss / dt, "unit": UNIT, "cores": 1, "sweeps_per_sec": sweeps / dt, "grid_steps_per_sec": gs / dt,
                "sample": "%d sweeps of cfg1 (main.py default: K=3, T=2, 19 observations, n~11), unmodified gauenk/smjp via oracle/ref_shim.py" % sweeps}
    except Exception as e:  # the port's line must still be printed
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference itself cannot
    travel to the GPU box) on all host cores, on a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    t_end = T_END
    sample_chains = procs  # one chain per worker per step
    for _ in range(args.warmup):
        cpu_sample(min(2, procs), 1, 10.0, min(2, procs))
    tot_ss = tot_gs = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ss, gs, _ = cpu_sample(sample_chains, SWEEPS_PER_REF_STEP, t_end, procs)
        tot_ss += ss
        tot_gs += gs
    wall = time.perf_counter() - t0
    val = tot_ss / wall
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(CHAINS_PER_GPU), "sample": "%d chains x %d sweeps per step" % (sample_chains, SWEEPS_PER_REF_STEP)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port",
                             "sample": "%d full-size cfg2 chains (T=100, n~500) x %d sweeps per step, NumPy oracle, "
                                       "%d processes" % (sample_chains, SWEEPS_PER_REF_STEP, procs)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            # beside the port: the unmodified Python reference itself, on the only configuration it finishes (cfg1)
            "reference_unmodified": time_unmodified_reference(),
            "chain_sweeps_per_sec": sample_chains * SWEEPS_PER_REF_STEP * args.steps / wall, "grid_steps_per_sec": tot_gs / wall}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
K10_CHAINS = 16384   # one GPU's shard of BASELINE configs[4] at 4 GPUs; north_star: ">= 4096 concurrent chains of a 10-state sMJP"
K10_T = 20.0
K10_K = 10


def k10_model():
    rng = np.random.default_rng(5)  # shapes ~ U(0.6, 3) (seed 5), scale 1, T = 20
    shape = rng.uniform(0.6, 3.0, (K10_K, K10_K))
    emis = np.ones((K10_K, K10_K)) + 9 * np.eye(K10_K)
    return shape, emis / emis.sum(axis=1, keepdims=True)


def k10_block(local, steps, peak, peak_src, chains=K10_CHAINS, chain_base=0, precisions=("f64", "f32"), opts=None):
    """The north_star target workload on one GPU: `chains` chains of the 10-state Weibull sMJP of cfg5 (T = 20,
    n ~ 206), device-resident Rao-Teh sweeps; per precision the sweep time, state-steps/s and the roofline of the FFBS
    kernel on RETAINED bytes."""
    import torch
    from smjp_b200.chains import DeviceChains
    from smjp_b200.engine import Engine
    shape, emis = k10_model()
    eng = Engine(local)
    eng.set_model(shape, np.ones((K10_K, K10_K)), OMEGA, emis, 1.0, K10_T)
    for k, v in (opts or {}).items():
        eng.set_option(k, v)
    obs_t = np.arange(OBS_DT, K10_T, OBS_DT)
    D = len(obs_t)
    out = {"workload": "cfg5 shard: %d chains, K=10 Weibull sMJP, shapes~U(0.6,3) seed 5, T=20, D=%d obs" % (chains, D)}
    names = {1: "ffbs_item_kernel", 5: "ffbs_fastk_kernel"}
    for prec in precisions:
        esz = 8 if prec == "f64" else 4
        ch = DeviceChains(eng, chains, obs_t, chain_base=chain_base)
        ch.init_prior(np.ones(K10_K) / K10_K, seed=SEED + 21).simulate_observations(seed=SEED + 22)
        ch.init_prior(np.ones(K10_K) / K10_K, seed=SEED + 23)
        for sw in range(3):
            ch.sweep(seed=SEED + 24, sweep=sw, precision=prec)
        torch.cuda.synchronize()
        eng.set_option("time_kernels", 1)
        eng.kernel_time()
        eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        snaps = []
        e0.record()
        for sw in range(steps):
            ch.sweep(seed=SEED + 24, sweep=3 + sw, precision=prec)
            snaps.append(ch.cur["offs"].clone())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        k_ms, k_cnt = eng.kernel_time()
        lp = float(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs"))
        eng.set_option("time_kernels", 0)
        ss = nominal = tail = gs = 0.0
        for offs in snaps:
            n = (offs[1:] - offs[:-1] - 1).cpu().numpy()
            ss += float(np.sum(K10_K * n.astype(np.float64) * (n + 1) / 2))
            gs += float(n.sum())
            nominal += algorithmic_bytes(n, D, esz, K10_K)
            tail += float(np.sum(40.0 * n + 12 * D))
        kid = int(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_kernel"))
        cap = ((ncu_traffic() or {}).get("k10_" + prec) or {})
        # the committed capture describes ffbs_fastk_kernel on this very workload (16384 chains); not other kernels / sizes
        on_captured_kernel = bool(cap) and names.get(kid) == "ffbs_fastk_kernel" and chains == K10_CHAINS
        out[prec] = {"value": ss / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
                     "chain_sweeps_per_sec": chains * steps / (ms / 1e3), "mean_grid_intervals": gs / (chains * steps),
                     "failed_chains": int((ch.status != 0).sum().item()),
                     "chains_redone_by_rescue_pass_last_launch": int(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_rescued")),
                     "roofline": roofline_block("%s<%s,10>" % (names.get(kid, "kernel %d" % kid), "double" if prec == "f64" else "float"),
                                                2 * esz * K10_K * lp + tail, nominal, k_cnt, k_ms, peak, peak_src,
                                                K10_K * lp / max(ss, 1.0), cap.get("dram_bytes_per_launch") if on_captured_kernel else None,
                                                compute=ceilings_from_capture(cap) if on_captured_kernel else None,
                                                kernel_share_of_step=k_ms / ms,
                                                grid=int(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_grid")),
                                                smem=int(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_smem")))}
        del ch
    eng.close()
    return out


PF_N = 65536      # BASELINE configs[2]: 65536 particles per chain, 5-state gamma-hazard sMJP, systematic resampling
PF_K = 5
PF_T = 10.0


def pf_block(rank, world, local, dev, runs_per_gpu, peak, peak_src, reps=2, family="gamma", opts=None):
    """BASELINE configs[2] as stated: bootstrap particle filter (pmcmc.py:224-310) with GAMMA holding-time hazards
    (shape ~ U(0.6, 3) seed 3, scale 1), K = 5, T = 10, D = 99 observations, 65536 particles per filter run, systematic
    resampling; `runs_per_gpu` filter runs in flight per GPU through the device-pointer entry point smjp_pf_dev.  With
    N > 1 GPUs every rank owns its runs and the log-normalisers are gathered over NCCL (gather_lognorm) inside the
    timed region.  Metric: particle-steps/s (N * D per run); roofline on SURVEY 8(d)'s 68 B per particle-step."""
    import torch
    import torch.distributed as dist
    from smjp_b200.distributed import gather_lognorm
    from smjp_b200.engine import Engine, pf_dev
    rng = np.random.default_rng(3)
    shape = rng.uniform(0.6, 3.0, (PF_K, PF_K))
    emis = np.ones((PF_K, PF_K)) + 9 * np.eye(PF_K)
    emis /= emis.sum(axis=1, keepdims=True)
    eng = Engine(local)
    eng.set_model(shape, np.ones((PF_K, PF_K)), OMEGA, emis, 1.0, PF_T)
    eng.set_option("chain_base", rank * runs_per_gpu)
    for k, v in (opts or {}).items():  # kernel experiments (tools/pf_probe.py)
        eng.set_option(k, v)
    obs_t = np.arange(OBS_DT, PF_T, OBS_DT)
    D = len(obs_t)
    y = rng.integers(0, PF_K, D)
    C = runs_per_gpu
    t = torch
    d_offs = t.arange(C + 1, dtype=t.int64, device=dev) * D
    d_ot = t.from_numpy(np.tile(obs_t, C)).to(dev)
    d_oy = t.from_numpy(np.tile(y, C).astype(np.int32)).to(dev)
    d_pi0 = t.full((PF_K,), 1.0 / PF_K, dtype=t.float64, device=dev)
    cap = 256
    d_po = t.arange(C + 1, dtype=t.int64, device=dev) * cap
    d_pv = t.zeros(C * cap, dtype=t.int32, device=dev)
    d_pt = t.zeros(C * cap, dtype=t.float64, device=dev)
    d_pl = t.zeros(C, dtype=t.int32, device=dev)
    d_st = t.zeros(C, dtype=t.int32, device=dev)
    d_ll = t.zeros(C, dtype=t.float64, device=dev)

    def run(seed):
        pf_dev(eng, C, d_offs, d_ot, d_oy, d_pi0, PF_N, seed, d_ll, d_po, d_pv, d_pt, d_pl, d_st, D, resampler="systematic",
               family=family)
        return gather_lognorm(d_ll) if world > 1 else d_ll

    run(SEED + 40)
    t.cuda.synchronize()
    eng.set_option("time_kernels", 1)
    eng.kernel_time("pf")
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    for r in range(reps):
        ll = run(SEED + 41 + r)
    e1.record()
    t.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    k_ms, k_cnt = eng.kernel_time("pf")
    eng.set_option("time_kernels", 0)
    bad = int((d_st != 0).sum().item())
    cs = int(eng.lib.smjp_ctx_get_stat(eng.ctx, b"pf_cluster"))
    if world > 1:
        st = t.tensor([ms, float(bad)], dtype=t.float64, device=dev)
        mx = st.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        ms, bad = float(mx[0]), int(mx[1])
    ll_h = ll.cpu().numpy()
    eng.close()
    steps_total = float(PF_N) * D * C * world * reps
    k_s = k_ms / max(k_cnt, 1) / 1e3
    ach = 68.0 * PF_N * D * C / k_s / 1e9
    return {"workload": "cfg3: particle filter, %d particles/run, K=5 %s hazards (shape~U(0.6,3) seed 3), T=10, D=%d, systematic "
                        "resampling, %d runs in flight per GPU" % (PF_N, family.upper(), D, C),
            "value": steps_total / (ms / 1e3), "unit": "particle-steps/s", "ms_per_call": ms / reps, "n_gpus": world,
            "failed_runs": bad, "cluster_size": cs,
            "loglik_mean": float(ll_h.mean()), "loglik_std_over_runs": float(ll_h.std()), "logliks_gathered": int(ll_h.shape[0]),
            "collective": "gather_lognorm: NCCL all_gather of [runs] log-normalisers, inside the timed region" if world > 1 else "none (1 GPU)",
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                         "kernel": "pf_kernel", "kernel_ms_avg": k_ms / max(k_cnt, 1), "algorithmic_bytes_per_launch": 68.0 * PF_N * D * C,
                         "bytes_per_particle_step": 68, "traffic": None}}


CFG5_CHAINS = 65536  # BASELINE configs[4]: 65536 chains, 10-state sMJP, SPLIT over the GPUs of the run


def cfg5_block(rank, world, local, dev, steps, barrier, overlap=False, interleave=1):
    """BASELINE configs[4] as stated: 65536 chains of the 10-state sMJP (k10_model, T = 20) split over the N ranks of
    this run (contiguous ranges, Philox streams keyed by the GLOBAL chain index), Rao-Teh sweeps with the NCCL
    all-gather of the posterior jump-count / holding-time summaries ([65536, 2K] doubles, mcmc_utils.py:40-56) once per
    sweep INSIDE the timed loop, on a side stream under the next sweep.  STRONG scaling: the total work is fixed.
    Rank 0 returns the block; `summaries_sha1` of the gathered summaries after the last sweep is the same at every N
    (sharded == unsharded, bit for bit).  interleave > 1: every rank runs its chains as that many sub-batches on their
    own streams (InterleavedChains): the candidate kernels of one sub-batch -- a tenth of an fp32 sweep at K = 10 -- run
    under the FFBS kernel of another.  Same chains, same bits: the SHA-1 does not change."""
    import hashlib
    import torch
    import torch.distributed as dist
    from smjp_b200.chains import DeviceChains, InterleavedChains
    from smjp_b200.distributed import SummaryGather, shard_range
    from smjp_b200.engine import Engine
    shape, emis = k10_model()
    lo, hi = shard_range(CFG5_CHAINS, rank, world)
    C = hi - lo
    obs_t = np.arange(OBS_DT, K10_T, OBS_DT)
    out = {"workload": "cfg5: %d chains total, K=10 Weibull sMJP, T=20, D=%d obs, split over %d GPU(s)" % (CFG5_CHAINS, len(obs_t), world),
           "scaling": "strong", "n_gpus": world, "chains_per_gpu": C, "unit": UNIT, "sub_batches_per_gpu": int(interleave)}
    model = (shape, np.ones((K10_K, K10_K)), OMEGA, emis, 1.0, K10_T)
    eng = Engine(local)
    eng.set_model(*model)
    cur_stream = torch.cuda.current_stream(dev)
    for prec in ("f64", "f32"):
        if interleave > 1:
            ch = InterleavedChains(local, C, obs_t, model, parts=interleave, chain_base=lo)
        else:
            ch = DeviceChains(eng, C, obs_t, chain_base=lo)
        ch.init_prior(np.ones(K10_K) / K10_K, seed=SEED + 31).simulate_observations(seed=SEED + 32)
        ch.init_prior(np.ones(K10_K) / K10_K, seed=SEED + 33)
        gather = SummaryGather(CFG5_CHAINS, K10_K, rank, world, dev, overlap=overlap)
        release = [None]

        def one_sweep(sw):
            ch.sweep(seed=SEED + 34, sweep=sw, precision=prec)
            if interleave > 1:
                t_, j_ = ch.summaries(release[0])
                gather.launch(t_, j_)
                release[0] = torch.cuda.Event()
                release[0].record(cur_stream)
                return ch.grid_offsets()
            gather.launch(ch.time_in_state, ch.jumps)
            return [ch.cur["offs"].clone()]

        for sw in range(3):
            one_sweep(sw)
        gather.result()
        gather.timing.clear()
        if interleave > 1:
            ch.join()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        snaps = []
        e0.record()
        if interleave > 1:
            ch.fork()
        for sw in range(steps):
            snaps += one_sweep(3 + sw)
        tis, jm = gather.result()
        if interleave > 1:
            ch.join()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        ss = gs = 0.0
        for offs in snaps:
            n = (offs[1:] - offs[:-1] - 1).to(torch.float64)
            ss += float((K10_K * n * (n + 1) / 2).sum().item())
            gs += float(n.sum().item())
        st = torch.tensor([ms, ss, gs, float((ch.status != 0).sum().item())], dtype=torch.float64, device=dev)
        if world > 1:
            mx = st.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = st.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, ss, gs, bad = float(mx[0]), float(sm[1]), float(sm[2]), int(sm[3])
        else:
            bad = int(st[3])
        h = hashlib.sha1(tis.cpu().numpy().tobytes() + jm.cpu().numpy().tobytes()).hexdigest()
        out[prec] = {"value": ss / (ms / 1e3), "ms_per_step": ms / steps, "steps": steps,
                     "chain_sweeps_per_sec": CFG5_CHAINS * steps / (ms / 1e3), "mean_grid_intervals": gs / (CFG5_CHAINS * steps),
                     "failed_chains": bad,
                     "collective": {"op": "all_gather_into_tensor [65536, 20] f64" if world > 1 else "none (1 GPU: staging copy only)",
                                    "backend": "nccl" if world > 1 else None, "per_sweep": True, "inside_timed_region": True,
                                    "bytes_per_gather": gather.bytes_per_gather, "gather_ms_avg": gather.gather_ms()},
                     "summaries_sha1": h}
        if interleave > 1:
            ch.close()
        del ch, gather
    eng.close()
    return out if rank == 0 else None


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from smjp_b200.chains import DeviceChains
    from smjp_b200.engine import Engine, sweep_host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # One small CTA per collective (1 channel, 64 threads).  The sweep kernels are persistent and -- with two
        # sub-batches interleaved -- keep every SM slot taken all the time; NCCL's default all-gather kernel wants most of
        # an SM's registers per CTA and then starves until the whole pipeline drains (measured at N=2: 10.84 ms per step
        # against 10.23 with this setting, 10.22 without any gather).  A 64-thread CTA fits the slot ONE exiting warp
        # frees.  The gathers move 0.4-10 MB per sweep on a side stream: their bandwidth is irrelevant.
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "1")
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")
        os.environ.setdefault("NCCL_NTHREADS", "64")
        # the collectives run on a HIGH-PRIORITY stream: the sweep kernels are persistent and keep every SM slot taken, so
        # the CTAs of an all-gather enqueued behind them would otherwise wait for a whole kernel to drain
        pg_opts = None
        try:
            pg_opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=pg_opts)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    C = args.chains
    prec = args.precision
    esz = 8 if prec == "f64" else 4

    eng = Engine(local)
    from smjp_b200.engine import _ptr  # noqa: F401
    emis = np.ones((K, K)) + 9 * np.eye(K)
    emis /= emis.sum(axis=1, keepdims=True)  # 10:1 emission, smjp_utils.py:552-557
    eng.set_model(SHAPE, np.ones((K, K)), OMEGA, emis, 1.0, T_END)
    eng.set_option("ffbs_threads", args.threads)
    eng.set_option("item_k3", args.item_k3)
    eng.set_option("window_log2", args.window_log2)
    eng.set_option("ffbs_warp", args.warp)
    eng.set_option("ffbs_ring", args.ring)
    eng.set_option("ffbs_fast", args.fast)
    eng.set_option("fast_ring", args.fast_ring)
    eng.set_option("fast_ways", args.fast_ways)
    eng.set_option("fast_cpb", args.fast_cpb)
    eng.set_option("speculate", 0 if args.no_speculate else 1)
    obs_t = np.arange(OBS_DT, T_END, OBS_DT)
    D = len(obs_t)
    # ground-truth path -> observations; independent prior draw as the initial state
    ch = DeviceChains(eng, C, obs_t, chain_base=rank * C)
    ch.init_prior(np.ones(K) / K, seed=SEED + 1).simulate_observations(seed=SEED + 2)
    ch.init_prior(np.ones(K) / K, seed=SEED + 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from smjp_b200.distributed import SummaryGather
    # posterior summaries of ALL chains on every rank, every sweep, inside the timed region: one NCCL all-gather of
    # [C, 2K] doubles per sweep on a side stream, overlapped with the next sweep (mcmc_utils.py:40-56 per chain).
    # A warm-up step is a timed step: it gathers too (NCCL sets its channels up on the first call of a communicator)
    gather = SummaryGather(C * world, K, rank, world, dev, overlap=bool(args.gather_overlap))
    for sw in range(args.warmup):
        ch.sweep(seed=SEED, sweep=sw, precision=prec)
        gather.launch(ch.time_in_state, ch.jumps)
    gather.result()
    gather.reset_counters()
    # the collector is kept out of the timed regions: a generation-2 pass of a process with torch loaded was
    # seen to add 5-50 ms to a single step
    gc.collect()
    gc.disable()
    barrier()
    eng.set_option("time_kernels", 1)
    eng.kernel_time()
    eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs")  # reset the live-window counter
    launches0 = eng.launch_count
    spec0 = [eng.lib.smjp_ctx_get_stat(eng.ctx, k) for k in (b"spec_launches", b"spec_misses")]
    sampler = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ss = gs = 0.0
    alg_bytes = tail_bytes = 0.0
    n_snap = []
    barrier()
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)] if os.environ.get("SMJP_BENCH_DEBUG") else None
    sampler.begin()
    ev0.record()
    for i in range(args.steps):
        ch.sweep(seed=SEED, sweep=args.warmup + i, precision=prec)
        gather.launch(ch.time_in_state, ch.jumps)
        n_snap.append(ch.cur["offs"].clone())  # this sweep's grid windows (the buffer is recycled two sweeps on)
        if step_ev:
            step_ev[i].record()
    all_tis, all_jumps = gather.result()  # waits for the last gather: inside the timed region
    ev1.record()
    barrier()
    sampler.end()
    ms = ev0.elapsed_time(ev1)
    gather_ms = gather.gather_ms()
    if step_ev and rank == 0:
        print("per-step ms:", [round(ev0.elapsed_time(e), 2) for e in step_ev], file=sys.stderr)
    launches = eng.launch_count - launches0
    spec = [eng.lib.smjp_ctx_get_stat(eng.ctx, k) - v0 for k, v0 in zip((b"spec_launches", b"spec_misses"), spec0)]
    k_ms, k_cnt = eng.kernel_time()
    live_pairs = float(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs"))
    eng.set_option("time_kernels", 0)
    for offs in n_snap:
        n = (offs[1:] - offs[:-1] - 1).cpu().numpy()
        ss += float(np.sum(K * n.astype(np.float64) * (n + 1) / 2))
        gs += float(n.sum())
        alg_bytes += algorithmic_bytes(n, D, esz)
        tail_bytes += float(np.sum(40.0 * n + 12 * D))
    status_bad = int((ch.status != 0).sum().item())

    # ---- end to end through the host-buffer C-ABI call (smjp_sweep_host) ----
    # host (pinned) buffers in, host buffers out, every step: H2D of the current paths and the
    # observations, D2H of the new grids, paths and metrics
    from smjp_b200.engine import PackedSweep
    V, T = ch.paths_host()
    Y = ch.obs_y.cpu().numpy()
    e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 10))
    ps = PackedSweep(eng, C, ch.obs_offs_h, np.tile(obs_t, C), Y, int(sum(len(v) for v in V) * 3))
    ps.set_paths(V, T)
    e_ss = 0.0
    h2d = d2h = 0
    e_wall = 1e-9
    e2e_parts = 1
    if e2e_steps and args.interleave > 1:
        # Two half-batches in flight through smjp_sweep_host_begin / _end, one engine context and stream each: while
        # one half's kernels run, the other half's inputs cross PCIe and its kernels queue up behind (and fill the
        # tail of the running persistent kernel).  Same chains (chain_base), same host arrays, every byte still
        # crosses PCIe inside the timed region.
        from smjp_b200.engine import Engine
        e2e_parts = args.interleave
        halves = []
        for s_ in range(e2e_parts):
            lo, hi = s_ * C // e2e_parts, (s_ + 1) * C // e2e_parts
            st_ = torch.cuda.Stream(dev)
            eh = Engine(local, stream=st_.cuda_stream)
            eh.set_model(SHAPE, np.ones((K, K)), OMEGA, emis, 1.0, T_END)
            for k_, v_ in (("ffbs_threads", args.threads), ("window_log2", args.window_log2), ("ffbs_warp", args.warp),
                           ("ffbs_fast", args.fast), ("speculate", 0 if args.no_speculate else 1), ("chain_base", rank * C + lo)):
                eh.set_option(k_, v_)
            ph = PackedSweep(eh, hi - lo, ch.obs_offs_h[lo: hi + 1] - ch.obs_offs_h[lo], np.tile(obs_t, hi - lo),
                             Y[ch.obs_offs_h[lo]: ch.obs_offs_h[hi]], int(sum(len(v) for v in V[lo:hi]) * 3))
            ph.set_paths(V[lo:hi], T[lo:hi])
            halves.append((st_, eh, ph))
        for i in range(2):  # warm the staging buffers
            for _, _, ph in halves:
                ph.run(SEED + 7, i, precision=prec).swap()
        barrier()
        t0 = time.perf_counter()
        for _, _, ph in halves:
            h2d += ph.h2d_bytes()
            ph.begin(SEED + 7, 2, precision=prec)
        for i in range(e2e_steps):
            for _, _, ph in halves:
                ph.end()
                d2h += ph.d2h_bytes()
                n = np.diff(ph.w_offs[: ph.C + 1]).astype(np.float64) - 1
                e_ss += float(np.sum(K * n * (n + 1) / 2))
                assert int(np.count_nonzero(ph.status[: ph.C])) == 0
                ph.swap()
                if i + 1 < e2e_steps:
                    h2d += ph.h2d_bytes()
                    ph.begin(SEED + 7, 3 + i, precision=prec)
        torch.cuda.synchronize()
        e_wall = max(time.perf_counter() - t0, 1e-9)
        e2e_in_place = {"observations": bool(halves[0][1].lib.smjp_ctx_get_stat(halves[0][1].ctx, b"obs_in_place")),
                        "paths_gathered": bool(halves[0][1].lib.smjp_ctx_get_stat(halves[0][1].ctx, b"paths_gathered"))}
        for _, eh, _ in halves:
            eh.close()
    elif e2e_steps:
        for i in range(2):  # warm the staging buffers
            ps.run(SEED + 7, i, precision=prec).swap()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            h2d += ps.h2d_bytes()
            ps.run(SEED + 7, 2 + i, precision=prec)
            d2h += ps.d2h_bytes()
            n = np.diff(ps.w_offs[: C + 1]).astype(np.float64) - 1
            e_ss += float(np.sum(K * n * (n + 1) / 2))
            ps.swap()
        torch.cuda.synchronize()
        e_wall = max(time.perf_counter() - t0, 1e-9)
        assert int(np.count_nonzero(ps.status[:C])) == 0
    if e2e_parts == 1:
        e2e_in_place = {"observations": bool(eng.lib.smjp_ctx_get_stat(eng.ctx, b"obs_in_place")),
                        "paths_gathered": bool(eng.lib.smjp_ctx_get_stat(eng.ctx, b"paths_gathered"))}

    # ---- the headline pass: the same chains as `--interleave` sub-batches on their own streams ----
    # The sweep kernels are persistent (a warp pulls chains until the queue is empty), so a launch ends in a tail during
    # which SMs drain; chains are independent, so sweep s+1 of one half of the batch can fill the tail of sweep s of the
    # other half.  Same chains, same global chain indices, same bits (tests/test_sweep_gpu.py); the pass above -- one
    # batch, one stream, every kernel alone on the GPU -- is what the roofline block is measured on.
    il = None
    il_ms = il_ss = il_gs = 0.0
    il_launches = il_bad = 0
    il_clk = None
    il_gather_ms = 0.0
    il_checksum = 0.0
    if args.interleave > 1:
        from smjp_b200.chains import InterleavedChains
        opts = {"ffbs_threads": args.threads, "item_k3": args.item_k3, "window_log2": args.window_log2, "ffbs_warp": args.warp,
                "ffbs_ring": args.ring, "ffbs_fast": args.fast, "fast_ring": args.fast_ring, "fast_ways": args.fast_ways,
                "fast_cpb": args.fast_cpb, "speculate": 0 if args.no_speculate else 1}
        il = InterleavedChains(local, C, obs_t, (SHAPE, np.ones((K, K)), OMEGA, emis, 1.0, T_END), parts=args.interleave,
                               chain_base=rank * C, options=opts)
        il.init_prior(np.ones(K) / K, seed=SEED + 1).simulate_observations(seed=SEED + 2)
        il.init_prior(np.ones(K) / K, seed=SEED + 3)
        gather2 = SummaryGather(C * world, K, rank, world, dev, overlap=bool(args.gather_overlap))
        cur_stream = torch.cuda.current_stream(dev)
        release = None
        for sw in range(args.warmup):
            il.sweep(seed=SEED, sweep=sw, precision=prec)
            tis2, jm2 = il.summaries(release)
            gather2.launch(tis2, jm2)
            release = torch.cuda.Event()
            release.record(cur_stream)
        gather2.result()
        gather2.reset_counters()
        il.join()
        gc.collect()
        barrier()
        launches_il0 = il.launch_count
        sampler2 = ClockSampler(local)
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        il_snaps = []
        barrier()
        sampler2.begin()
        i0.record(cur_stream)
        il.fork()
        for i in range(args.steps):
            il.sweep(seed=SEED, sweep=args.warmup + i, precision=prec)
            il_snaps.append(il.grid_offsets())
            tis2, jm2 = il.summaries(release)
            gather2.launch(tis2, jm2)
            release = torch.cuda.Event()
            release.record(cur_stream)
        all_tis2, _ = gather2.result()  # waits for the last gather: inside the timed region
        il.join()
        i1.record(cur_stream)
        barrier()
        sampler2.end()
        il_ms = i0.elapsed_time(i1)
        il_gather_ms = gather2.gather_ms()
        il_clk = sampler2.summary()
        il_launches = il.launch_count - launches_il0
        for offs_parts in il_snaps:
            for offs in offs_parts:
                n = (offs[1:] - offs[:-1] - 1).cpu().numpy().astype(np.float64)
                il_ss += float(np.sum(K * n * (n + 1) / 2))
                il_gs += float(n.sum())
        il_bad = int((il.status != 0).sum().item())
        il_checksum = float(all_tis2.sum().item())
        il_rows = int(all_tis2.shape[0])
        il_bytes_per_gather = gather2.bytes_per_gather
        il.close()

    peak, peak_src = measured_peak()
    pf = None
    if not args.no_pf:
        # as BASELINE states it (gamma holding times: incomplete-gamma inversion per draw), and the same filter with the
        # reference's own Weibull hazards (distributions.py:296-304) for comparison
        # "many runs in flight" = one filter run per SM (a run is one cluster; at this size one CTA of 1024 threads)
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        pf = {"gamma_runs_8": pf_block(rank, world, local, dev, 8, peak, peak_src),
              "gamma_runs_%d" % sms: pf_block(rank, world, local, dev, sms, peak, peak_src, reps=1),
              "weibull_runs_8": pf_block(rank, world, local, dev, 8, peak, peak_src, reps=3, family="weibull"),
              "weibull_runs_%d" % sms: pf_block(rank, world, local, dev, sms, peak, peak_src, reps=3, family="weibull")}
    cfg5 = None if args.no_cfg5 else cfg5_block(rank, world, local, dev, max(3, min(args.steps, 5)), barrier, bool(args.gather_overlap),
                                                   interleave=args.cfg5_interleave)

    # ---- reduce over ranks: time = max, work = sum ----
    ss_rank0 = ss  # this rank's own state-steps (the kernel figures below are rank 0's)
    stats = torch.tensor([ms, ss, gs, alg_bytes, float(launches), k_ms, float(k_cnt), e_wall, e_ss, float(status_bad),
                          float(h2d), float(d2h), il_ms, il_ss, il_gs, float(il_launches), float(il_bad)],
                         dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e_wall = float(mx[0]), float(mx[7])
        ss, gs, alg_total, launches, e_ss = float(sm[1]), float(sm[2]), float(sm[3]), float(sm[4]), float(sm[8])
        status_bad = int(sm[9])
        h2d, d2h = int(sm[10]), int(sm[11])
        il_ms = float(mx[12])
        il_ss, il_gs, il_launches, il_bad = float(sm[13]), float(sm[14]), float(sm[15]), int(sm[16])
    else:
        alg_total = alg_bytes
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    secs = ms / 1e3
    value = ss / secs
    single = {"value": value, "ms_per_step": ms / args.steps, "gpu_launches": int(launches), "failed_chains": status_bad,
              "what": "one batch on one stream: every kernel runs alone (the roofline block is measured on this pass)"}
    peak, peak_src = measured_peak()
    k_avg_s = (k_ms / max(k_cnt, 1)) / 1e3  # rank 0's FFBS kernel, average launch
    live_frac = live_pairs / max(ss_rank0 / K, 1.0)
    ret_bytes = 2 * esz * K * live_pairs + tail_bytes
    traffic = ncu_traffic()

    def compute_ceiling(precision, pairs_per_launch, kernel_s, sm_mhz):
        """second ceiling: the arithmetic pipe the hazard generation runs on.  Instructions per
        (i, j) pair and issue utilisation come from the committed ncu capture; the pipe fraction is live:
        pairs x thread-instructions per pair / (pipe lanes per clock x SMs x measured SM clock x kernel time)."""
        t = ((traffic or {}).get(precision) or {})
        if "pipe_thread_inst_per_pair" not in t or not sm_mhz:
            return None
        lanes = float(t["pipe_lanes_per_clk_sm"])
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        peak = lanes * sms * float(sm_mhz) * 1e6
        need = pairs_per_launch * float(t["pipe_thread_inst_per_pair"])
        return {"pipe": t["pipe"], "thread_inst_per_pair": t["pipe_thread_inst_per_pair"], "lanes_per_clk_sm": lanes,
                "peak_thread_inst_per_s": peak, "frac": need / (peak * kernel_s),
                "kernel_ms_at_pipe_peak": 1e3 * need / peak, "warp_inst_per_pair_all_pipes": t.get("warp_inst_per_warp_pair"),
                "issue_active_pct_ncu": t.get("issue_active_pct"), "source": "profiles/roofline_traffic.json (ncu --set full)"}

    clk = sampler.summary()
    head = {"value": value, "ms": ms, "gs": gs, "launches": launches, "bad": status_bad, "clk": clk}
    if il is not None:  # the headline: the interleaved pass
        head = {"value": il_ss / (il_ms / 1e3), "ms": il_ms, "gs": il_gs, "launches": il_launches, "bad": il_bad, "clk": il_clk}
        single["clocks"] = clk
    kernel_main = FFBS_KERNEL_NAMES.get(int(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_kernel")), "ffbs_item_kernel")
    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": prec, "data": "synthetic",
        "config": {"workload": workload_name(C), "chains_total": C * world,
                   "parallelism": "chains sharded x%d" % world + (
                       "; per GPU %d interleaved sub-batches on their own streams (the next kernel of one fills the tail of the "
                       "other's persistent kernel; same chains, same bits)" % args.interleave if il is not None else ""),
                   "l2": "no flush needed: the message stream of one sweep (%.1f GB/GPU nominal, 15 GB of measured DRAM traffic in fp64 "
                         "with the live window) exceeds the 126 MB L2 a hundredfold" % (alg_bytes / args.steps / 1e9),
                   "mean_grid_intervals": gs / (C * world * args.steps), "candidates": "reference (Q5)"},
        "chain_sweeps_per_sec": C * world * args.steps / (head["ms"] / 1e3), "grid_steps_per_sec": head["gs"] / (head["ms"] / 1e3),
        "gpu_launches": int(head["launches"]), "failed_chains": head["bad"],
        "single_stream": single,
        "rescue": {"chains_redone_last_launch": int(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_rescued")),
                   "ring_boost": int(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_fast_boost_f64"))},
        "clocks": head["clk"],
        "roofline": roofline_block(
            "%s<%s,3>" % (kernel_main, "double" if prec == "f64" else "float"), ret_bytes, alg_bytes, k_cnt, k_ms, peak, peak_src,
            live_frac, captured_traffic(traffic, prec, kernel_main),
            grid=eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_grid"), threads=eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_threads"),
            smem=eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_smem"), kernel_share_of_step=k_ms / ms,
            measured_on="the single-stream pass (`single_stream`): the kernel timed alone on the GPU, CUDA events around every "
                        "launch; in the interleaved headline pass two launches overlap in time",
            # the kernel skips grid points whose states have underflowed to EXACT zeros (results unchanged to the
            # last bit): `achieved`/`frac` count the states actually retained, `nominal_*` all K n(n+1)/2
            window="exact (zeros only)" if args.window_log2 == 0 else "pruned below 2^%d of the row sum" % args.window_log2,
            compute=compute_ceiling(prec, ss_rank0 / K / max(k_cnt, 1), k_avg_s, clk.get("sm_mhz"))),
        # sweeps whose kernels were enqueued for the grid size predicted from the previous sweep, before their own
        # counts reached the host, and how many of those had to be launched again (rank 0)
        "speculative_launches": {"launches": spec[0], "repeated": spec[1]},
        # the collective of the data path: posterior summaries of every chain gathered on every rank once per sweep,
        # inside the timed region, overlapped with the next sweep on a side stream
        "collective": {"op": "all_gather_into_tensor [chains_total, 2K] f64 (time in state, jump counts)" if world > 1 else "none (1 GPU: staging copy only)",
                       "backend": "nccl" if world > 1 else None, "per_sweep": True, "inside_timed_region": True,
                       "overlapped_with_next_sweep": bool(args.gather_overlap),
                       "bytes_per_gather": gather.bytes_per_gather, "gather_ms_avg": il_gather_ms if il is not None else gather_ms,
                       "rows_gathered": int(all_tis.shape[0]), "time_in_state_checksum": il_checksum if il is not None else float(all_tis.sum().item()),
                       "time_in_state_checksum_single_stream": float(all_tis.sum().item())},
        "cfg5": cfg5,
        "pf": pf,
        "e2e": {"value": e_ss / e_wall, "unit": UNIT, "h2d_bytes_per_step": h2d // max(e2e_steps, 1),
                "d2h_bytes_per_step": d2h // max(e2e_steps, 1), "steps": e2e_steps,
                "batches_in_flight": e2e_parts,
                "api": ("smjp_sweep_host_begin / _end, two half-batches in flight on two contexts" if e2e_parts > 1 else "smjp_sweep_host") +
                       " (pinned host buffers in/out, all PCIe traffic inside the timed region: the paths' entries in use are pulled by a gather kernel, each chain's observation slice is read in place by the FFBS kernel's bulk copies (TMA) when the chain starts, the grid is copied back under the FFBS kernel, new paths / metrics / status are written by the kernels straight into the pinned output buffers)",
                "inputs_read_in_place": e2e_in_place},
    }
    if world == 1 and prec == "f64" and not args.no_alt:
        # the same workload in the engine's fp32 mode (north_star: messages within 1e-4), reported beside the
        # fp64 headline: device-resident sweeps only
        for sw in range(3):
            ch.sweep(seed=SEED + 11, sweep=sw, precision="f32")
        torch.cuda.synchronize()
        eng.set_option("time_kernels", 1)
        eng.kernel_time()
        eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs")
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        snaps = []
        a0.record()
        dbg = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)] if os.environ.get("SMJP_BENCH_DEBUG") else None
        for sw in range(args.steps):
            ch.sweep(seed=SEED + 11, sweep=3 + sw, precision="f32")
            snaps.append(ch.cur["offs"].clone())
            if dbg:
                dbg[sw].record()
        a1.record()
        torch.cuda.synchronize()
        ms32 = a0.elapsed_time(a1)
        if dbg:
            print("f32 per-step ms:", [round(a0.elapsed_time(e), 2) for e in dbg], file=sys.stderr)
        k32, c32 = eng.kernel_time()
        lp32 = float(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs"))
        eng.set_option("time_kernels", 0)
        ss32 = b32 = r32 = 0.0
        for offs in snaps:
            n = (offs[1:] - offs[:-1] - 1).cpu().numpy()
            ss32 += float(np.sum(K * n.astype(np.float64) * (n + 1) / 2))
            b32 += algorithmic_bytes(n, D, 4)
            r32 += float(np.sum(40.0 * n + 12 * D))
        r32 += 2 * 4 * K * lp32
        kernel32 = FFBS_KERNEL_NAMES.get(int(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_kernel")), "ffbs_item_kernel")
        line["f32"] = {"value": ss32 / (ms32 / 1e3), "unit": UNIT, "ms_per_step": ms32 / args.steps,
                       "chain_sweeps_per_sec": C * args.steps / (ms32 / 1e3),
                       "failed_chains": int((ch.status != 0).sum().item()),
                       "roofline": roofline_block(
                           kernel32 + "<float,3>", r32, b32, c32, k32, peak, peak_src,
                           K * lp32 / max(ss32, 1.0), captured_traffic(traffic, "f32", kernel32),
                           compute=compute_ceiling("f32", ss32 / K / max(c32, 1), k32 / max(c32, 1) / 1e3, clk.get("sm_mhz"))),
                       "tolerance": "messages within 1e-4 of the reference (tests/test_ffbs_gpu.py)"}
    if world == 1 and prec == "f64" and not args.no_alt and args.window_log2 == 0:
        # the same fp64 workload with live-window PRUNING (SURVEY.md 7 hard part 2): states below 2^-70 of their
        # row's sum leave the window.  An approximation (far below fp64 rounding of the row sums), hence reported
        # beside the exact headline, never as it.
        eng.set_option("window_log2", -70)
        for sw in range(3):
            ch.sweep(seed=SEED + 13, sweep=sw, precision="f64")
        torch.cuda.synchronize()
        eng.set_option("time_kernels", 1)
        eng.kernel_time()
        eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs")
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        snaps = []
        p0.record()
        for sw in range(args.steps):
            ch.sweep(seed=SEED + 13, sweep=3 + sw, precision="f64")
            snaps.append(ch.cur["offs"].clone())
        p1.record()
        torch.cuda.synchronize()
        msp = p0.elapsed_time(p1)
        kp, cp = eng.kernel_time()
        lp = float(eng.lib.smjp_ctx_get_stat(eng.ctx, b"ffbs_live_pairs"))
        eng.set_option("time_kernels", 0)
        eng.set_option("window_log2", 0)
        ssp = 0.0
        for offs in snaps:
            n = (offs[1:] - offs[:-1] - 1).cpu().numpy()
            ssp += float(np.sum(K * n.astype(np.float64) * (n + 1) / 2))
        line["pruned"] = {"window": "states below 2^-70 of the row sum are dropped", "dtype": "f64",
                          "nominal_state_steps_per_sec": ssp / (msp / 1e3), "retained_state_steps_per_sec": K * lp / (msp / 1e3),
                          "live_fraction": K * lp / max(ssp, 1.0), "unit": UNIT, "ms_per_step": msp / args.steps,
                          "chain_sweeps_per_sec": C * args.steps / (msp / 1e3), "kernel_ms_avg": kp / max(cp, 1),
                          "failed_chains": int((ch.status != 0).sum().item())}
    if world == 1 and not args.no_alt:
        # north_star's stated target: >= 4096 concurrent chains of a 10-state sMJP (BASELINE configs[4]'s chains)
        line["k10"] = k10_block(local, max(3, min(args.steps, 6)), peak, peak_src)
    gc.enable()
    if world == 1 and not args.no_cpu:
        ss_c, gs_c, wall_c = cpu_sample(16, 3, T_END, 1)  # ~10-20 s of CPU work
        line["cpu_baseline"] = {"value": ss_c / wall_c, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "16 full-size cfg2 chains (T=100, n~500) x 3 sweeps, NumPy oracle "
                                          "(oracle/smjp_oracle.py), 1 process",
                                "chain_sweeps_per_sec": 48 / wall_c}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-alt", action="store_true", help="skip the secondary fp32 measurement")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (kernel experiments)")
    ap.add_argument("--gather-overlap", type=int, default=1, dest="gather_overlap",
                    help="1: the per-sweep NCCL gather runs on a side stream under the next sweep; 0: ordered between sweeps")
    ap.add_argument("--interleave", type=int, default=2,
                    help="headline pass: sub-batches per GPU, each on its own stream (1 = the single-stream pass is the headline)")
    ap.add_argument("--cfg5-interleave", type=int, default=1, dest="cfg5_interleave",
                    help="sub-batches per GPU in the cfg5 block (measured within +-2 %% of 1 at every N: the shards have 7-55 chains "
                         "per resident warp, the tail that interleaving hides is small, and at 8 GPUs half a shard no longer fills the GPU)")
    ap.add_argument("--no-pf", action="store_true", dest="no_pf", help="skip the cfg3 particle-filter block")
    ap.add_argument("--no-cfg5", action="store_true", dest="no_cfg5", help="skip the cfg5 strong-scaling block")
    ap.add_argument("--threads", type=int, default=0, help="threads per CTA of the FFBS kernel (0 = auto)")
    ap.add_argument("--item-k3", type=int, default=0, help="experiment: item-parallel kernel for K=3")
    ap.add_argument("--warp", type=int, default=-1, help="warp-per-chain ring kernel (ffbs_warp.cuh): -1 auto, 0 off, 1 on")
    ap.add_argument("--no-speculate", action="store_true", dest="no_speculate",
                    help="wait for every sweep's grid sizes on the host before launching its FFBS kernel")
    ap.add_argument("--fast", type=int, default=-1, help="ffbs_fast.cuh (deferred normalisation, split barriers): -1 auto, 0 off, 1 on")
    ap.add_argument("--fast-cpb", type=int, default=0, dest="fast_cpb", help="fast kernel: chains per CTA, one warp each (0 auto)")
    ap.add_argument("--fast-ways", type=int, default=0, dest="fast_ways", help="fast kernel: grid pairs in flight per thread (0 auto, 1, 2, 4)")
    ap.add_argument("--fast-ring", type=int, default=0, dest="fast_ring", help="its ring slots (0 = auto)")
    ap.add_argument("--ring", type=int, default=0, help="ring slots of 32 grid points for --warp (0 = auto)")
    ap.add_argument("--window-log2", type=int, default=0, dest="window_log2",
                    help="0: exact live window (default); -L: prune states below 2^-L of the row sum")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
