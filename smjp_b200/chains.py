"""Device-resident batch of independent chains: the batched Rao-Teh sweep the reference lacks.

State per chain is its current trajectory (V, T) (raoteh.py:25, :58-59), kept in HBM as ragged
arrays; one `sweep()` = candidates -> forward filter -> backward sample -> compaction -> metrics for
every chain, entirely on the device (smjp_sweep_dev).  torch is used for device memory only.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check
from .engine import PRECISIONS, Engine, _ptr, ragged_offsets

CAND_MODES = {"reference": _lib.CAND_REFERENCE, "exact": _lib.CAND_EXACT}


class DeviceChains:
    def __init__(self, engine, C, obs_t, device=None, chain_base=0):
        """obs_t: one 1-D array of observation times shared by all chains, or a list per chain."""
        import torch
        self.torch = torch
        self.eng = engine
        self.C = int(C)
        self.dev = torch.device("cuda", engine.device) if device is None else device
        self.chain_base = int(chain_base)
        engine.set_option("chain_base", self.chain_base)
        if isinstance(obs_t, np.ndarray) and obs_t.ndim == 1 and obs_t.dtype != object:
            obs_t = [obs_t] * C
        self.obs_offs_h = ragged_offsets([len(t) for t in obs_t])
        flat = np.concatenate([np.asarray(t, dtype=np.float64) for t in obs_t]) if self.obs_offs_h[-1] else np.zeros(0)
        self.obs_offs = torch.from_numpy(self.obs_offs_h).to(self.dev)
        self.obs_t = torch.from_numpy(flat).to(self.dev)
        self.obs_y = torch.zeros(len(flat), dtype=torch.int32, device=self.dev)
        self.K = engine.K
        self.time_in_state = torch.zeros((C, self.K), dtype=torch.float64, device=self.dev)
        self.jumps = torch.zeros((C, self.K), dtype=torch.int32, device=self.dev)
        self.status = torch.zeros(C, dtype=torch.int32, device=self.dev)
        self.cur = None   # dict(offs, v, t, len) of the current paths
        self.nxt = None
        self.W = None
        self.soff = None
        self.n_max = 0
        self.total_w = 0
        self.sweeps_done = 0
        self.cap = 0

    # ------------------------------------------------------------------------------------------
    def _alloc_paths(self, capacity, C=None):
        t = self.torch
        C = self.C if C is None else C
        return {"offs": t.zeros(C + 1, dtype=t.int64, device=self.dev),
                "v": t.zeros(capacity, dtype=t.int32, device=self.dev),
                "t": t.zeros(capacity, dtype=t.float64, device=self.dev),
                "len": t.zeros(C, dtype=t.int32, device=self.dev), "cap": capacity}

    def init_prior(self, pi0, seed, honour_scale=False, cap_per_chain=None):
        """sample_smjp_trajectory_prior for every chain (raoteh.py:25)."""
        t = self.torch
        pi0 = np.ascontiguousarray(pi0, dtype=np.float64)
        self.eng.set_option("chain_base", self.chain_base)  # the engine is shared: another DeviceChains may have set its own
        cap = cap_per_chain or 64
        while True:
            p = self._alloc_paths(cap * self.C)
            p["offs"].copy_(t.arange(self.C + 1, dtype=t.int64, device=self.dev) * cap)
            check(self.eng.lib.smjp_prior_dev(self.eng.ctx, self.C, _ptr(p["offs"]), _ptr(pi0), int(honour_scale),
                                              int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(p["v"]), _ptr(p["t"]),
                                              _ptr(p["len"]), _ptr(self.status)))
            if int((self.status != 0).sum().item()) == 0:
                break
            cap *= 2
            if cap > (1 << 22):
                raise _lib.EngineError("prior path does not terminate")
        self.cur = p
        return self

    def set_paths(self, V, T):
        """host lists of per-chain (state indices, times) -> device state"""
        t = self.torch
        offs = ragged_offsets([len(v) for v in V])
        p = self._alloc_paths(int(offs[-1]))
        p["offs"].copy_(t.from_numpy(offs))
        p["v"].copy_(t.from_numpy(np.concatenate([np.asarray(v, dtype=np.int32) for v in V])))
        p["t"].copy_(t.from_numpy(np.concatenate([np.asarray(x, dtype=np.float64) for x in T])))
        p["len"].copy_(t.from_numpy(np.diff(offs).astype(np.int32)))
        self.cur = p
        return self

    def simulate_observations(self, seed):
        """y_d ~ emis[v(t_d)] from the current paths (sample_data_posterior, smjp_utils.py:350-364)."""
        p = self.cur
        self.eng.set_option("chain_base", self.chain_base)
        check(self.eng.lib.smjp_simulate_obs_dev(self.eng.ctx, self.C, _ptr(p["offs"]), _ptr(p["v"]), _ptr(p["t"]),
                                                 _ptr(p["len"]), _ptr(self.obs_offs), _ptr(self.obs_t),
                                                 int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(self.obs_y)))
        return self

    def set_observations(self, obs_y):
        t = self.torch
        if isinstance(obs_y, np.ndarray) and obs_y.ndim == 1 and obs_y.dtype != object:
            obs_y = [obs_y] * self.C
        flat = np.concatenate([np.asarray(y, dtype=np.int32) for y in obs_y]) if self.obs_offs_h[-1] else np.zeros(0, np.int32)
        if len(flat) != len(self.obs_y):
            raise ValueError("observation count mismatch")
        self.obs_y.copy_(t.from_numpy(flat))
        return self

    # ------------------------------------------------------------------------------------------
    def sweep(self, seed, sweep=None, mode="reference", precision="f64", dense=False):
        """One Gibbs sweep of every chain (raoteh.py:58-59).  Asynchronous after one size read-back.
        dense=True runs the K-state collapse (shape == 1 models only; K <= 64, the limit of smjp_set_model)."""
        t = self.torch
        sweep = self.sweeps_done if sweep is None else sweep
        self.eng.set_option("chain_base", self.chain_base)
        cur = self.cur
        if self.cap == 0:  # first sweep: size from the live path entries, grow on demand
            self.cap = int(cur["len"].sum().item()) * 3 + 1024
        n_max = ctypes.c_int64(0)
        total = ctypes.c_int64(0)
        while True:
            if self.nxt is None or self.nxt["cap"] < self.cap:
                self.nxt = self._alloc_paths(self.cap)
            if self.W is None or self.W.numel() < self.cap:
                self.W = t.zeros(self.cap, dtype=t.float64, device=self.dev)
            if self.soff is None or self.soff.numel() < cur["cap"]:
                self.soff = t.zeros(cur["cap"], dtype=t.int32, device=self.dev)
            nx = self.nxt
            # the new paths and the grid share one bound: a recycled path buffer may be larger than W
            cap_call = min(int(nx["cap"]), int(self.W.numel()))
            rc = (self.eng.lib.smjp_sweep_dense_dev if dense else self.eng.lib.smjp_sweep_dev)(
                self.eng.ctx, self.C, _ptr(cur["offs"]), _ptr(cur["v"]), _ptr(cur["t"]), _ptr(cur["len"]),
                _ptr(self.obs_offs), _ptr(self.obs_t), _ptr(self.obs_y), int(seed) & 0xFFFFFFFFFFFFFFFF, int(sweep),
                CAND_MODES[mode], PRECISIONS[precision], cap_call, _ptr(nx["offs"]), _ptr(self.soff),
                _ptr(self.W), _ptr(nx["v"]), _ptr(nx["t"]), _ptr(nx["len"]), _ptr(self.time_in_state),
                _ptr(self.jumps), _ptr(self.status), ctypes.byref(n_max), ctypes.byref(total))
            if rc == -3 and total.value > cap_call:  # buffers too small for this sweep's grids: regrow, redo
                self.cap = int(total.value * 1.5) + 1024
                self.nxt = None
                self.W = None
                continue
            check(rc)
            break
        self.n_max, self.total_w = int(n_max.value), int(total.value)
        self.cur, self.nxt = nx, (cur if cur["cap"] >= self.cap else None)
        self.sweeps_done += 1
        return self

    # ------------------------------------------------------------------------------------------
    def shape_mle(self, pooled=True, grid=0.001, grid_max=5.0):
        """Parameter-step statistics of the current trajectories (raoteh.py:119-167) on the device: for every
        (v, v') the number of holding times, sum ln T, sum T and the reference's grid-search Weibull shape
        estimate, per chain (pooled=False -> arrays [C,K,K]) or over all chains (pooled=True -> [K,K]).
        khat / g are NaN where the transition never occurs."""
        t = self.torch
        G = 1 if pooled else self.C
        K = self.K
        khat = t.empty((G, K, K), dtype=t.float64, device=self.dev)
        g = t.empty_like(khat)
        sl = t.empty_like(khat)
        st = t.empty_like(khat)
        cnt = t.empty((G, K, K), dtype=t.int32, device=self.dev)
        p = self.cur
        check(self.eng.lib.smjp_shape_mle_dev(self.eng.ctx, self.C, K, _ptr(p["offs"]), _ptr(p["v"]), _ptr(p["t"]),
                                              _ptr(p["len"]), int(bool(pooled)), float(grid), float(grid_max),
                                              _ptr(khat), _ptr(g), _ptr(cnt), _ptr(sl), _ptr(st)))
        self.eng.sync()
        out = {"khat": khat.cpu().numpy(), "g": g.cpu().numpy(), "count": cnt.cpu().numpy(),
               "sum_log": sl.cpu().numpy(), "sum_t": st.cpu().numpy()}
        return {k: (v[0] if pooled else v) for k, v in out.items()}

    # ------------------------------------------------------------------------------------------
    def record(self, history):
        """append this sweep's per-chain metrics to a summaries.SummaryHistory (stays on the device)"""
        history.append(self.time_in_state, self.jumps)
        return self

    def paths_host(self):
        """current trajectories as host lists (V state indices, T times)"""
        p = self.cur
        offs = p["offs"].cpu().numpy()
        lens = p["len"].cpu().numpy()
        v = p["v"].cpu().numpy()
        tt = p["t"].cpu().numpy()
        V = [v[offs[c]: offs[c] + lens[c]].copy() for c in range(self.C)]
        T = [tt[offs[c]: offs[c] + lens[c]].copy() for c in range(self.C)]
        return V, T

    def grids_host(self):
        """grids W of the last sweep"""
        offs = self.cur["offs"].cpu().numpy()
        w = self.W[: self.total_w].cpu().numpy()
        return [w[offs[c]: offs[c + 1]].copy() for c in range(self.C)]

    def state_steps_last_sweep(self):
        """live augmented state-steps K*n(n+1)/2 summed over chains (SURVEY.md 8d unit)"""
        n = (self.cur["offs"][1:] - self.cur["offs"][:-1] - 1).to(self.torch.float64)
        return float((self.K * n * (n + 1) / 2).sum().item()), float(n.sum().item())


class InterleavedChains:
    """The same batch of chains as `parts` sub-batches, each with its own engine context and CUDA stream.

    Why: the sweep kernels are persistent -- a CTA pulls chains until the queue is empty -- so a launch ends in a tail
    during which SMs drain (cfg2: 4096 chains over 1184 resident warps, ~8 % of the warp slots idle on average).  The
    chains of a Gibbs sampler are independent, so sweep s+1 of one half of the batch needs nothing from sweep s of the
    other half: with two sub-batches in flight on two streams the block scheduler fills the slots one kernel frees
    with CTAs of the other's next kernel, and only the last launch of a run has a tail.  Chain c keeps its global
    index (`chain_base`), so every path is bit-identical to the single-batch run (Philox streams are keyed by chain).

    sweep() enqueues one sweep of every sub-batch and returns; join() makes the caller's stream wait for all of them.
    The per-chain results (time_in_state, jumps, status) are gathered in global chain order by summaries()."""

    def __init__(self, device_index, C, obs_t, model, parts=2, chain_base=0, options=None):
        """model: (shape, scale, omega, emis, power, t_end) as for Engine.set_model; options: {name: value} for every engine"""
        import torch
        from .engine import Engine
        self.torch = torch
        self.C, self.S = int(C), int(parts)
        self.dev = torch.device("cuda", int(device_index))
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(self.S)]
        self.bounds = [(s * self.C // self.S, (s + 1) * self.C // self.S) for s in range(self.S)]
        self.engines, self.parts = [], []
        for s, (lo, hi) in enumerate(self.bounds):
            with torch.cuda.stream(self.streams[s]):
                eng = Engine(int(device_index), stream=self.streams[s].cuda_stream)
                eng.set_model(*model)
                for k, v in (options or {}).items():
                    eng.set_option(k, v)
                per = obs_t if (isinstance(obs_t, np.ndarray) and obs_t.ndim == 1 and obs_t.dtype != object) else list(obs_t)[lo:hi]
                self.engines.append(eng)
                self.parts.append(DeviceChains(eng, hi - lo, per, device=self.dev, chain_base=int(chain_base) + lo))
        self.K = self.engines[0].K
        self._tis = torch.zeros((2, self.C, self.K), dtype=torch.float64, device=self.dev)   # double-buffered snapshots
        self._jumps = torch.zeros((2, self.C, self.K), dtype=torch.int32, device=self.dev)
        self._snap = 0
        self._released = None  # event: the previous snapshot buffer of this parity may be overwritten

    def _each(self, fn):
        for s, ch in enumerate(self.parts):
            with self.torch.cuda.stream(self.streams[s]):
                fn(ch)
        return self

    def set_option(self, key, value):
        for e in self.engines:
            e.set_option(key, value)
        return self

    def init_prior(self, pi0, seed, **kw):
        return self._each(lambda ch: ch.init_prior(pi0, seed, **kw))

    def simulate_observations(self, seed):
        return self._each(lambda ch: ch.simulate_observations(seed))

    def sweep(self, seed, sweep=None, mode="reference", precision="f64"):
        """one Gibbs sweep of every chain: one enqueue per sub-batch, each on its own stream"""
        return self._each(lambda ch: ch.sweep(seed, sweep=sweep, mode=mode, precision=precision))

    def join(self, stream=None):
        """the given (default: current) stream waits for everything enqueued on the sub-batch streams"""
        t = self.torch
        stream = t.cuda.current_stream(self.dev) if stream is None else stream
        for s in self.streams:
            ev = t.cuda.Event()
            ev.record(s)
            stream.wait_event(ev)
        return self

    def fork(self, stream=None):
        """the sub-batch streams wait for what is enqueued on the given (default: current) stream"""
        t = self.torch
        stream = t.cuda.current_stream(self.dev) if stream is None else stream
        ev = t.cuda.Event()
        ev.record(stream)
        for s in self.streams:
            s.wait_event(ev)
        return self

    def summaries(self, release_after=None):
        """(time_in_state [C, K] f64, jumps [C, K] i32) of the last sweep in global chain order: every sub-batch copies
        its rows on its own stream into one of two snapshot buffers and the CURRENT stream waits for the copies --
        no host synchronisation.  A buffer is reused two calls later; pass as `release_after` an event after which the
        buffer handed out by the PREVIOUS call may be overwritten (e.g. recorded once its all-gather was staged)."""
        t = self.torch
        b = self._snap
        self._snap ^= 1
        if release_after is not None:
            for s in self.streams:
                s.wait_event(release_after)
        cur = t.cuda.current_stream(self.dev)
        for s, (ch, (lo, hi)) in enumerate(zip(self.parts, self.bounds)):
            with t.cuda.stream(self.streams[s]):
                self._tis[b, lo:hi].copy_(ch.time_in_state)
                self._jumps[b, lo:hi].copy_(ch.jumps)
                ev = t.cuda.Event()
                ev.record(self.streams[s])
            cur.wait_event(ev)
        return self._tis[b], self._jumps[b]

    @property
    def status(self):
        self.join()
        return self.torch.cat([ch.status for ch in self.parts])

    def grid_offsets(self):
        """per sub-batch: the grid windows (offs [C_part + 1]) of the last sweep -- clones, on the sub-batch streams"""
        out = []
        for s, ch in enumerate(self.parts):
            with self.torch.cuda.stream(self.streams[s]):
                out.append(ch.cur["offs"].clone())
        return out

    def paths_host(self):
        self.join()
        self.torch.cuda.current_stream(self.dev).synchronize()
        V, T = [], []
        for ch in self.parts:
            v, w = ch.paths_host()
            V += v
            T += w
        return V, T

    @property
    def launch_count(self):
        return sum(e.launch_count for e in self.engines)

    def close(self):
        self.join()
        self.torch.cuda.current_stream(self.dev).synchronize()
        for e in self.engines:
            e.close()
