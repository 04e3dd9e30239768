"""Batched engine object: owns one C-ABI context (device + stream) and exposes the kernels on ragged
batches of chains, with HOST (NumPy) buffers or DEVICE (torch) tensors.  This is the layer the
reference-named drop-in functions (smjp_b200.smjp_utils, .fast_hmm, .raoteh, .pmcmc) call into.
PyTorch is used only for device memory and streams."""
import ctypes

import numpy as np

from . import _lib
from ._lib import SMJP_F32, SMJP_F64, check

PRECISIONS = {"f64": SMJP_F64, "fp64": SMJP_F64, "float64": SMJP_F64, np.float64: SMJP_F64,
              "f32": SMJP_F32, "fp32": SMJP_F32, "float32": SMJP_F32, np.float32: SMJP_F32}


def _ptr(a):
    """address of a NumPy array / torch tensor / None"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


def ragged_offsets(lengths):
    offs = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lengths, dtype=np.int64), out=offs[1:])
    return offs


def pack_ragged(arrays, dtype):
    """list of 1-D arrays -> (flat, offsets)"""
    lens = [len(a) for a in arrays]
    offs = ragged_offsets(lens)
    flat = np.zeros(int(offs[-1]), dtype=dtype)
    for a, o in zip(arrays, offs[:-1]):
        flat[o:o + len(a)] = a
    return flat, offs


def message_offsets(n_per_chain, K):
    """element offsets of each chain's triangular message block: K*n*(n+1)/2 elements per chain"""
    n = np.asarray(n_per_chain, dtype=np.int64)
    return ragged_offsets(K * n * (n + 1) // 2)


def triangular_to_dense(tri, n, K):
    """one chain's exported messages -> the reference's dense [n, K*n] array (fast_hmm.py:209;
    column a = v*n + j, zeros where the reference holds log 0)."""
    dense = np.zeros((n, K, n), dtype=tri.dtype)
    pos = 0
    for i in range(n):
        dense[i, :, : i + 1] = tri[pos: pos + K * (i + 1)].reshape(K, i + 1)
        pos += K * (i + 1)
    return dense.reshape(n, K * n)


class Engine:
    def __init__(self, device=0, stream=None):
        self.lib = _lib.load()
        self.device = int(device)
        self.ctx = self.lib.smjp_ctx_create(self.device, stream)
        if not self.ctx:
            raise _lib.EngineError(self.lib.smjp_last_error().decode())
        self.K = None
        self.Kobs = None
        # experiment switches for whole-suite runs: SMJP_FFBS_WARP=1 (warp-per-chain ring kernel), SMJP_FFBS_RING=R
        import os
        for env, key in (("SMJP_FFBS_WARP", "ffbs_warp"), ("SMJP_FFBS_RING", "ffbs_ring"), ("SMJP_FFBS_FAST", "ffbs_fast")):
            if os.environ.get(env):
                self.set_option(key, int(os.environ[env]))

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.smjp_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def set_option(self, key, value):
        check(self.lib.smjp_ctx_set_option(self.ctx, key.encode(), int(value)))

    def sync(self):
        check(self.lib.smjp_ctx_sync(self.ctx))

    @property
    def launch_count(self):
        return int(self.lib.smjp_ctx_launch_count(self.ctx))

    def kernel_time(self, kernel="ffbs"):
        """(total ms, launches) of the named kernel since the last call (needs option time_kernels=1)"""
        ms, cnt = ctypes.c_double(0), ctypes.c_int64(0)
        check(self.lib.smjp_ctx_kernel_time(self.ctx, kernel.encode(), ctypes.byref(ms), ctypes.byref(cnt)))
        return ms.value, cnt.value

    def set_model(self, shape, scale, omega, emis, likelihood_power, t_end, obs_combine="sum"):
        shape = np.ascontiguousarray(shape, dtype=np.float64)
        scale = np.ascontiguousarray(scale, dtype=np.float64)
        emis = np.ascontiguousarray(emis, dtype=np.float64)
        if shape.ndim != 2 or shape.shape[0] != shape.shape[1] or scale.shape != shape.shape:
            raise ValueError("shape/scale must be [K,K]")
        K = shape.shape[0]
        if emis.ndim != 2 or emis.shape[0] != K:
            raise ValueError("emis must be [K,Kobs]")
        check(self.lib.smjp_set_model(self.ctx, K, emis.shape[1], _ptr(shape), _ptr(scale), float(omega),
                                      _ptr(emis), float(likelihood_power), float(t_end)))
        self.set_option("obs_product", {"sum": 0, "product": 1}[obs_combine])
        self.K, self.Kobs = K, emis.shape[1]
        self.omega = float(omega)
        self.shape, self.scale, self.emis = shape.copy(), scale.copy(), emis.copy()
        self.t_end = float(t_end)

    # ------------------------------------------------------------------------------------------
    def ffbs_host(self, W, obs_t, obs_y, uniforms=None, seed=0, sweep=0, precision="f64",
                  want_messages=False, want_logZ=True, want_aug_idx=True, want_metrics=True):
        """FFBS on HOST data.  W: list (per chain) of grids; obs_t/obs_y: list per chain, or one
        array shared by all chains.  Returns a dict of NumPy results; ragged outputs as lists."""
        if self.K is None:
            raise ValueError("set_model first")
        C = len(W)
        prec = PRECISIONS[precision]
        Wf, w_offs = pack_ragged([np.asarray(w, dtype=np.float64) for w in W], np.float64)
        shared = (isinstance(obs_t, np.ndarray) and obs_t.dtype != object and obs_t.ndim == 1) or \
                 (len(obs_t) == 0 or np.isscalar(obs_t[0]))
        if shared:  # one observation record for every chain
            obs_t = [obs_t] * C
            obs_y = [obs_y] * C
        if len(obs_t) != C or len(obs_y) != C:
            raise ValueError("need one observation record per chain")
        of, o_offs = pack_ragged([np.asarray(t, dtype=np.float64) for t in obs_t], np.float64)
        oy, _ = pack_ragged([np.asarray(y, dtype=np.int32) for y in obs_y], np.int32)
        if len(oy) and (oy.min() < 0 or oy.max() >= self.Kobs):
            raise ValueError("observation symbol outside [0, Kobs)")
        uf = None
        if uniforms is not None:
            uf = np.zeros(len(Wf), dtype=np.float64)
            for c in range(C):
                u = np.asarray(uniforms[c], dtype=np.float64)
                n = w_offs[c + 1] - w_offs[c] - 1
                if len(u) != n:
                    raise ValueError("chain %d: need one uniform per grid interval (%d), got %d" % (c, n, len(u)))
                uf[w_offs[c]: w_offs[c] + n] = u
        n_per = np.diff(w_offs) - 1
        K = self.K
        rdt = np.float64 if prec == SMJP_F64 else np.float32
        msgs = m_offs = None
        if want_messages:
            m_offs = message_offsets(n_per, K)
            msgs = np.zeros(int(m_offs[-1]), dtype=rdt)
        total = len(Wf)
        logZ = np.zeros(total, dtype=np.float64) if want_logZ else None
        aug = np.zeros(total, dtype=np.int32) if want_aug_idx else None
        pv = np.zeros(total, dtype=np.int32)
        pt = np.zeros(total, dtype=np.float64)
        pl = np.zeros(C, dtype=np.int32)
        tis = np.zeros((C, K), dtype=np.float64) if want_metrics else None
        jm = np.zeros((C, K), dtype=np.int32) if want_metrics else None
        st = np.zeros(C, dtype=np.int32)
        check(self.lib.smjp_ffbs_host(self.ctx, C, _ptr(w_offs), _ptr(Wf), _ptr(o_offs), _ptr(of), _ptr(oy),
                                      _ptr(uf), int(seed) & 0xFFFFFFFFFFFFFFFF, int(sweep), prec,
                                      _ptr(msgs), _ptr(m_offs), _ptr(logZ), _ptr(aug), _ptr(pv), _ptr(pt),
                                      _ptr(pl), _ptr(tis), _ptr(jm), _ptr(st)))
        out = {"status": st, "path_len": pl, "n": n_per, "w_offs": w_offs}
        out["V"] = [pv[w_offs[c]: w_offs[c] + pl[c]].copy() for c in range(C)]
        out["T"] = [pt[w_offs[c]: w_offs[c] + pl[c]].copy() for c in range(C)]
        if want_logZ:
            out["logZ"] = [logZ[w_offs[c]: w_offs[c] + n_per[c]].copy() for c in range(C)]
        if want_aug_idx:
            out["aug_idx"] = [aug[w_offs[c]: w_offs[c] + n_per[c]].copy() for c in range(C)]
        if want_metrics:
            out["time_in_state"], out["jumps"] = tis, jm
        if want_messages:
            out["messages"] = [msgs[m_offs[c]: m_offs[c + 1]] for c in range(C)]
        return out

    def ffbs_dense(self, W, obs_t, obs_y, u_state=None, u_jump=None, seed=0, sweep=0, precision="f64"):
        """FFBS of a shape == 1 model on its K-state collapse (smjp_ffbs_dense_dev; K <= 64).
        Host lists in, host lists out (staged through torch device tensors); uniforms per grid interval."""
        import torch as t
        if self.K is None:
            raise ValueError("set_model first")
        C = len(W)
        dev = t.device("cuda", self.device)
        Wf, w_offs = pack_ragged([np.asarray(w, dtype=np.float64) for w in W], np.float64)
        if (isinstance(obs_t, np.ndarray) and obs_t.dtype != object and obs_t.ndim == 1) or len(obs_t) == 0 or np.isscalar(obs_t[0]):
            obs_t, obs_y = [obs_t] * C, [obs_y] * C
        of, o_offs = pack_ragged([np.asarray(x, dtype=np.float64) for x in obs_t], np.float64)
        oy, _ = pack_ragged([np.asarray(y, dtype=np.int32) for y in obs_y], np.int32)
        if len(oy) and (oy.min() < 0 or oy.max() >= self.Kobs):
            raise ValueError("observation symbol outside [0, Kobs)")

        def flat(us):
            if us is None:
                return None
            f = np.zeros(len(Wf), dtype=np.float64)
            for c in range(C):
                n = w_offs[c + 1] - w_offs[c] - 1
                u = np.asarray(us[c], dtype=np.float64)
                if len(u) != n:
                    raise ValueError("chain %d: need one uniform per grid interval (%d), got %d" % (c, n, len(u)))
                f[w_offs[c]: w_offs[c] + n] = u
            return t.from_numpy(f).to(dev)

        d = lambda a: t.from_numpy(a).to(dev)
        total = len(Wf)
        dW, dwo, doo, dot, doy = d(Wf), d(w_offs), d(o_offs), d(of), d(oy)
        dus, duj = flat(u_state), flat(u_jump)
        logZ = t.zeros(total, dtype=t.float64, device=dev)
        pv = t.zeros(total, dtype=t.int32, device=dev)
        pt = t.zeros(total, dtype=t.float64, device=dev)
        pl = t.zeros(C, dtype=t.int32, device=dev)
        tis = t.zeros((C, self.K), dtype=t.float64, device=dev)
        jm = t.zeros((C, self.K), dtype=t.int32, device=dev)
        st = t.zeros(C, dtype=t.int32, device=dev)
        t.cuda.synchronize(dev)
        check(self.lib.smjp_ffbs_dense_dev(self.ctx, C, _ptr(dwo), _ptr(dW), _ptr(doo), _ptr(dot), _ptr(doy),
                                           _ptr(dus), _ptr(duj), int(seed) & 0xFFFFFFFFFFFFFFFF, int(sweep),
                                           PRECISIONS[precision], _ptr(logZ), _ptr(pv), _ptr(pt), _ptr(pl),
                                           _ptr(tis), _ptr(jm), _ptr(st), total))
        self.sync()
        pl_h, pv_h, pt_h, lz = pl.cpu().numpy(), pv.cpu().numpy(), pt.cpu().numpy(), logZ.cpu().numpy()
        n_per = np.diff(w_offs) - 1
        return {"status": st.cpu().numpy(), "path_len": pl_h, "n": n_per,
                "V": [pv_h[w_offs[c]: w_offs[c] + pl_h[c]].copy() for c in range(C)],
                "T": [pt_h[w_offs[c]: w_offs[c] + pl_h[c]].copy() for c in range(C)],
                "logZ": [lz[w_offs[c]: w_offs[c] + n_per[c]].copy() for c in range(C)],
                "time_in_state": tis.cpu().numpy(), "jumps": jm.cpu().numpy()}

    def ffbs_dev(self, C, w_offs, W, obs_offs, obs_t, obs_y, path_v, path_t, path_len, status, n_max,
                 total_w, uniforms=None, seed=0, sweep=0, precision="f64", messages=None, msg_offs=None,
                 logZ=None, aug_idx=None, time_in_state=None, jumps=None):
        """FFBS on DEVICE tensors (torch, already resident); asynchronous on the context's stream."""
        check(self.lib.smjp_ffbs_dev(self.ctx, int(C), _ptr(w_offs), _ptr(W), _ptr(obs_offs), _ptr(obs_t),
                                     _ptr(obs_y), _ptr(uniforms), int(seed) & 0xFFFFFFFFFFFFFFFF, int(sweep),
                                     PRECISIONS[precision], _ptr(messages), _ptr(msg_offs), _ptr(logZ),
                                     _ptr(aug_idx), _ptr(path_v), _ptr(path_t), _ptr(path_len),
                                     _ptr(time_in_state), _ptr(jumps), _ptr(status), int(n_max), int(total_w)))


def pf_dev(eng, C, obs_offs, obs_t, obs_y, pi0, N, seed, loglik, p_offs, path_v, path_t, path_len, status,
           max_obs_per_chain, resampler="systematic", family="weibull", logZ=None, extend=False):
    """Bootstrap particle filter (smjp_smc, pmcmc.py:224-310) on DEVICE tensors (torch), asynchronous on the context's
    stream: nothing is staged and nothing synchronises, so pmcmc over many chains stays resident."""
    res = {"multinomial": _lib.RESAMPLE_MULTINOMIAL, "systematic": _lib.RESAMPLE_SYSTEMATIC}[resampler]
    eng.set_option("pf_extend", int(bool(extend)))
    eng.set_option("hazard_family", {"weibull": 0, "gamma": 1}[family])
    try:
        check(eng.lib.smjp_pf_dev(eng.ctx, int(C), _ptr(obs_offs), _ptr(obs_t), _ptr(obs_y), _ptr(pi0), int(N),
                                  int(seed) & 0xFFFFFFFFFFFFFFFF, res, int(max_obs_per_chain), _ptr(logZ), _ptr(loglik),
                                  _ptr(p_offs), _ptr(path_v), _ptr(path_t), _ptr(path_len), _ptr(status)))
    finally:
        eng.set_option("hazard_family", 0)


def _pack_paths(V, T):
    lens = [len(v) for v in V]
    if any(len(t) != l for t, l in zip(T, lens)):
        raise ValueError("V and T lengths differ")
    offs = ragged_offsets(lens)
    pv = np.concatenate([np.asarray(v, dtype=np.int32) for v in V]) if lens else np.zeros(0, np.int32)
    pt = np.concatenate([np.asarray(t, dtype=np.float64) for t in T]) if lens else np.zeros(0)
    return offs, np.ascontiguousarray(pv), np.ascontiguousarray(pt), np.asarray(lens, dtype=np.int32)


def candidates_host(eng, V, T, seed, sweep=0, mode="reference"):
    """Thinned-event grids for HOST paths (sample_smjp_event_times, smjp_utils.py:194-210).
    V, T: lists per chain (state indices, times).  Returns the list of grids W."""
    C = len(V)
    offs, pv, pt, pl = _pack_paths(V, T)
    w_offs = np.zeros(C + 1, dtype=np.int64)
    total = ctypes.c_int64(0)
    cap = int(len(pv) * max(2.0, float(getattr(eng, "omega", 2.0)) + 1) + 64)
    modes = {"reference": _lib.CAND_REFERENCE, "exact": _lib.CAND_EXACT}
    while True:
        W = np.zeros(cap, dtype=np.float64)
        rc = eng.lib.smjp_candidates_host(eng.ctx, C, _ptr(offs), _ptr(pv), _ptr(pt), _ptr(pl),
                                          int(seed) & 0xFFFFFFFFFFFFFFFF, int(sweep), modes[mode], _ptr(w_offs),
                                          _ptr(W), cap, ctypes.byref(total))
        if rc == -3 and total.value > cap:
            cap = int(total.value)
            continue
        check(rc)
        break
    return [W[w_offs[c]: w_offs[c + 1]].copy() for c in range(C)]


class PackedSweep:
    """Reusable HOST buffers (pinned through torch when available) for smjp_sweep_host: the packed form
    of one Rao-Teh sweep for C chains, with no per-chain Python work.  `inputs` are (offs, v, t, len) of
    the current paths; after `run()` the outputs (w_offs, W, v, t, len, time_in_state, jumps, status)
    hold the new grids and paths and `swap()` makes them the next inputs."""

    def __init__(self, eng, C, obs_offs, obs_t, obs_y, capacity):
        self.eng, self.C, self.K = eng, int(C), eng.K
        self.obs_offs = self._buf(np.int64, C + 1, obs_offs)
        self.obs_t = self._buf(np.float64, len(obs_t), obs_t)
        self.obs_y = self._buf(np.int32, len(obs_y), obs_y)
        self._alloc(int(capacity))
        self.total_w = 0

    @staticmethod
    def _buf(dtype, n, init=None):
        try:
            import torch
            t = torch.empty(max(int(n), 1), dtype={np.int64: torch.int64, np.int32: torch.int32, np.float64: torch.float64}[dtype],
                            pin_memory=torch.cuda.is_available())
            a = t.numpy()  # the view keeps the pinned tensor alive through .base
        except Exception:
            a = np.empty(max(int(n), 1), dtype=dtype)
        if init is not None:
            a[: len(init)] = init
        return a

    def _alloc(self, capacity):
        C, K = self.C, self.K
        self.cap = capacity
        self.in_offs, self.in_len = self._buf(np.int64, C + 1), self._buf(np.int32, C)
        self.in_v, self.in_t = self._buf(np.int32, capacity), self._buf(np.float64, capacity)
        self.w_offs, self.out_len = self._buf(np.int64, C + 1), self._buf(np.int32, C)
        self.W = self._buf(np.float64, capacity)
        self.out_v, self.out_t = self._buf(np.int32, capacity), self._buf(np.float64, capacity)
        self.tis, self.jumps = self._buf(np.float64, C * K), self._buf(np.int32, C * K)
        self.status = self._buf(np.int32, C)

    def set_paths(self, V, T):
        offs, pv, pt, pl = _pack_paths(V, T)
        if len(pv) > self.cap:
            self._alloc(int(len(pv) * 3 + 1024))
        self.in_offs[: self.C + 1] = offs
        self.in_v[: len(pv)] = pv
        self.in_t[: len(pt)] = pt
        self.in_len[: self.C] = pl
        return self

    def run(self, seed, sweep, mode="reference", precision="f64"):
        """one sweep, synchronous: the outputs are valid on return"""
        return self.begin(seed, sweep, mode, precision).end()

    def end(self):
        """wait for the sweep enqueued by begin(); the outputs are valid afterwards"""
        check(self.eng.lib.smjp_sweep_host_end(self.eng.ctx))
        return self

    def begin(self, seed, sweep, mode="reference", precision="f64"):
        """enqueue one sweep (smjp_sweep_host_begin) and return with its kernels and copies in flight: a caller with two
        PackedSweeps (two engines) alternates begin / end so that one batch is staged while the other computes"""
        modes = {"reference": _lib.CAND_REFERENCE, "exact": _lib.CAND_EXACT}
        total = ctypes.c_int64(0)
        while True:
            rc = self.eng.lib.smjp_sweep_host_begin(
                self.eng.ctx, self.C, _ptr(self.in_offs), _ptr(self.in_v), _ptr(self.in_t), _ptr(self.in_len),
                _ptr(self.obs_offs), _ptr(self.obs_t), _ptr(self.obs_y), int(seed) & 0xFFFFFFFFFFFFFFFF, int(sweep),
                modes[mode], PRECISIONS[precision], self.cap, _ptr(self.w_offs), _ptr(self.W), _ptr(self.out_v),
                _ptr(self.out_t), _ptr(self.out_len), _ptr(self.tis), _ptr(self.jumps), _ptr(self.status),
                ctypes.byref(total))
            if rc == -3 and total.value > self.cap:
                keep = (self.in_offs.copy(), self.in_v.copy(), self.in_t.copy(), self.in_len.copy())
                self._alloc(int(total.value * 1.5) + 1024)
                self.in_offs[: len(keep[0])] = keep[0]
                self.in_v[: len(keep[1])] = keep[1]
                self.in_t[: len(keep[2])] = keep[2]
                self.in_len[: len(keep[3])] = keep[3]
                continue
            check(rc)
            break
        self.total_w = int(total.value)
        return self

    def swap(self):
        """outputs of the last sweep become the inputs of the next (paths live in the grid windows)"""
        self.in_offs, self.w_offs = self.w_offs, self.in_offs
        self.in_v, self.out_v = self.out_v, self.in_v
        self.in_t, self.out_t = self.out_t, self.in_t
        self.in_len, self.out_len = self.out_len, self.in_len
        return self

    def h2d_bytes(self):
        """bytes that cross PCIe towards the device in the next run(): path entries (the windows as a whole when they
        are copied, the entries in use when the library's gather kernel pulls them out of these pinned buffers -- stat
        `paths_gathered` of the previous run), path lengths and offsets, the observations (uploaded, or read in place
        by the FFBS kernel's bulk copies: same bytes either way)"""
        gathered = bool(self.eng.lib.smjp_ctx_get_stat(self.eng.ctx, b"paths_gathered"))
        live = int(self.in_len[: self.C].sum()) if gathered else int(self.in_offs[self.C])
        return live * 12 + self.C * 12 + 8 + (self.C + 1) * 8 + len(self.obs_t) * 12

    def d2h_bytes(self):
        """bytes that reached the host in the last run(): the grid (copied while the FFBS kernel runs), the grid
        offsets, and -- written by the kernels straight into these pinned buffers -- the live path entries,
        path lengths, status words and per-state metrics"""
        live = int(self.out_len[: self.C].sum())
        return self.total_w * 8 + live * 12 + (self.C + 1) * 8 + self.C * 8 + self.C * self.K * 12


def sweep_host(eng, V, T, obs_t, obs_y, seed, sweep=0, mode="reference", precision="f64"):
    """One Rao-Teh sweep (raoteh.py:58-59) for HOST paths; returns dict(W, V, T, time_in_state, jumps, status)."""
    C = len(V)
    offs, pv, pt, pl = _pack_paths(V, T)
    if isinstance(obs_t, np.ndarray) and obs_t.ndim == 1 and obs_t.dtype != object:
        obs_t, obs_y = [obs_t] * C, [obs_y] * C
    of, o_offs = pack_ragged([np.asarray(t, dtype=np.float64) for t in obs_t], np.float64)
    oy, _ = pack_ragged([np.asarray(y, dtype=np.int32) for y in obs_y], np.int32)
    K = eng.K
    modes = {"reference": _lib.CAND_REFERENCE, "exact": _lib.CAND_EXACT}
    cap = int(len(pv) * 3 + 1024)
    total = ctypes.c_int64(0)
    while True:
        w_offs = np.zeros(C + 1, dtype=np.int64)
        W = np.zeros(cap, dtype=np.float64)
        nv = np.zeros(cap, dtype=np.int32)
        nt = np.zeros(cap, dtype=np.float64)
        nl = np.zeros(C, dtype=np.int32)
        tis = np.zeros((C, K), dtype=np.float64)
        jm = np.zeros((C, K), dtype=np.int32)
        st = np.zeros(C, dtype=np.int32)
        rc = eng.lib.smjp_sweep_host(eng.ctx, C, _ptr(offs), _ptr(pv), _ptr(pt), _ptr(pl), _ptr(o_offs), _ptr(of),
                                     _ptr(oy), int(seed) & 0xFFFFFFFFFFFFFFFF, int(sweep), modes[mode],
                                     PRECISIONS[precision], cap, _ptr(w_offs), _ptr(W), _ptr(nv), _ptr(nt), _ptr(nl),
                                     _ptr(tis), _ptr(jm), _ptr(st), ctypes.byref(total))
        if rc == -3 and total.value > cap:
            cap = int(total.value * 1.25) + 1024
            continue
        check(rc)
        break
    return {"W": [W[w_offs[c]: w_offs[c + 1]].copy() for c in range(C)],
            "V": [nv[w_offs[c]: w_offs[c] + nl[c]].copy() for c in range(C)],
            "T": [nt[w_offs[c]: w_offs[c] + nl[c]].copy() for c in range(C)],
            "time_in_state": tis, "jumps": jm, "status": st}


def hmm_dense_host(eng, P, pi0, E=None, emis=None, y=None, uniforms=None, seed=0, precision="f64", sample=True):
    """Dense S-state HMM for a batch of chains (hidden_markov_model.py:213-313, :88-202).
    P: [S,S] or [n,S,S]; per chain either E[c]: [n_c,S] likelihoods, or symbols y[c]: [n_c] with the
    table emis [S,Kobs].  Returns dict(alpha_hat, logc, loglik, states, status) with ragged lists."""
    P = np.ascontiguousarray(P, dtype=np.float64)
    S = P.shape[-1]
    p_steps = 0 if P.ndim == 2 else P.shape[0]
    pi0 = np.ascontiguousarray(pi0, dtype=np.float64)
    prec = PRECISIONS[precision]
    if E is not None:
        C = len(E)
        lens = [len(e) for e in E]
        Ef = np.ascontiguousarray(np.concatenate([np.asarray(e, dtype=np.float64).reshape(-1, S) for e in E]))
        em, yf, Kobs = None, None, 0
    else:
        C = len(y)
        lens = [len(v) for v in y]
        Ef = None
        em = np.ascontiguousarray(emis, dtype=np.float64)
        Kobs = em.shape[1]
        yf = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.int32) for v in y]))
    offs = ragged_offsets(lens)
    total = int(offs[-1])
    uf = None
    if uniforms is not None:
        uf = np.ascontiguousarray(np.concatenate([np.asarray(u, dtype=np.float64) for u in uniforms]))
        if len(uf) != total:
            raise ValueError("need one uniform per step")
    rdt = np.float64 if prec == SMJP_F64 else np.float32
    ah = np.zeros((total, S), dtype=rdt)
    logc = np.zeros(total, dtype=np.float64)
    ll = np.zeros(C, dtype=np.float64)
    st = np.zeros(C, dtype=np.int32)
    states = np.zeros(total, dtype=np.int32) if sample else None
    check(eng.lib.smjp_hmm_dense_host(eng.ctx, C, S, _ptr(offs), _ptr(P), p_steps, _ptr(Ef), _ptr(em), Kobs, _ptr(yf),
                                      _ptr(pi0), _ptr(uf), int(seed) & 0xFFFFFFFFFFFFFFFF, prec, _ptr(ah), _ptr(logc),
                                      _ptr(ll), _ptr(states), _ptr(st)))
    sl = [slice(offs[c], offs[c + 1]) for c in range(C)]
    return {"alpha_hat": [ah[x] for x in sl], "logc": [logc[x] for x in sl], "loglik": ll, "status": st,
            "states": [states[x] for x in sl] if sample else None}


def pf_host(eng, obs_t, obs_y, pi0, N, seed, resampler="systematic", C=None, path_cap=None, extend=False,
            family="weibull"):
    """Bootstrap particle filter (smjp_smc, pmcmc.py:224-310) for C chains.  obs_t/obs_y: list per
    chain, or one shared record with C given.  Returns dict(V, T, loglik, logZ, status)."""
    if isinstance(obs_t, np.ndarray) and obs_t.ndim == 1 and obs_t.dtype != object:
        C = 1 if C is None else C
        obs_t, obs_y = [obs_t] * C, [obs_y] * C
    C = len(obs_t)
    of, o_offs = pack_ragged([np.asarray(t, dtype=np.float64) for t in obs_t], np.float64)
    oy, _ = pack_ragged([np.asarray(v, dtype=np.int32) for v in obs_y], np.int32)
    pi0 = np.ascontiguousarray(pi0, dtype=np.float64)
    res = {"multinomial": _lib.RESAMPLE_MULTINOMIAL, "systematic": _lib.RESAMPLE_SYSTEMATIC}[resampler]
    eng.set_option("pf_extend", int(bool(extend)))  # False: path held flat after the last observation (reference)
    eng.set_option("hazard_family", {"weibull": 0, "gamma": 1}[family])  # gamma: shape = a, scale = theta
    cap = path_cap or 256
    while True:
        p_offs = np.arange(C + 1, dtype=np.int64) * cap
        pv = np.zeros(C * cap, dtype=np.int32)
        pt = np.zeros(C * cap, dtype=np.float64)
        pl = np.zeros(C, dtype=np.int32)
        st = np.zeros(C, dtype=np.int32)
        ll = np.zeros(C, dtype=np.float64)
        lz = np.zeros(max(len(of), 1), dtype=np.float64)
        check(eng.lib.smjp_pf_host(eng.ctx, C, _ptr(o_offs), _ptr(of), _ptr(oy), _ptr(pi0), int(N),
                                   int(seed) & 0xFFFFFFFFFFFFFFFF, res, _ptr(lz), _ptr(ll), _ptr(p_offs), _ptr(pv),
                                   _ptr(pt), _ptr(pl), _ptr(st)))
        if np.any(st & _lib.ST_OVERFLOW) and cap < (1 << 20):
            cap *= 4
            continue
        break
    eng.set_option("hazard_family", 0)
    return {"V": [pv[c * cap: c * cap + pl[c]].copy() for c in range(C)],
            "T": [pt[c * cap: c * cap + pl[c]].copy() for c in range(C)],
            "loglik": ll, "logZ": [lz[o_offs[c]: o_offs[c + 1]].copy() for c in range(C)], "status": st}


def shape_mle_host(eng, V, T, K, pooled=True, grid=0.001, grid_max=5.0):
    """smjp_shape_mle_dev on HOST paths (lists of index arrays V and time arrays T): the parameter-step
    statistics of raoteh.py:119-167.  Returns dict(khat, g, count, sum_log, sum_t), [K,K] when pooled else [C,K,K]."""
    import torch as t
    C = len(V)
    dev = t.device("cuda", eng.device)
    offs, pv, pt, pl = _pack_paths(V, T)
    if len(pv) and (pv.min() < 0 or pv.max() >= K):
        raise ValueError("state index outside [0, K)")
    G = 1 if pooled else C
    d = lambda a: t.from_numpy(a).to(dev)
    dv, do, dt, dl = d(pv), d(offs), d(pt), d(pl)
    khat = t.empty((G, K, K), dtype=t.float64, device=dev)
    g, sl, st = t.empty_like(khat), t.empty_like(khat), t.empty_like(khat)
    cnt = t.empty((G, K, K), dtype=t.int32, device=dev)
    t.cuda.synchronize(dev)
    check(eng.lib.smjp_shape_mle_dev(eng.ctx, C, int(K), _ptr(do), _ptr(dv), _ptr(dt), _ptr(dl), int(bool(pooled)),
                                     float(grid), float(grid_max), _ptr(khat), _ptr(g), _ptr(cnt), _ptr(sl), _ptr(st)))
    eng.sync()
    out = {"khat": khat.cpu().numpy(), "g": g.cpu().numpy(), "count": cnt.cpu().numpy(),
           "sum_log": sl.cpu().numpy(), "sum_t": st.cpu().numpy()}
    return {k: (v[0] if pooled else v) for k, v in out.items()}


def raise_for_status(status):
    """Per-chain status bits -> the exceptions the reference raises at the same conditions."""
    status = np.asarray(status)
    if np.any(status & _lib.ST_BAD_INPUT):
        raise ValueError("observation symbol outside [0, Kobs)")
    if np.any(status & _lib.ST_GRID_NOT_INCREASING):
        raise ValueError("We can not have time_p >= time_c")  # smjp_utils.py:398-399
    if np.any(status & _lib.ST_DEGENERATE):
        raise FloatingPointError("forward message / sampling vector is zero or not finite "
                                 "(hidden_markov_model.py:170-185 exits here)")
    if np.any(status & _lib.ST_OVERFLOW):
        raise _lib.EngineError("engine capacity overflow (status %s)" % np.unique(status))
