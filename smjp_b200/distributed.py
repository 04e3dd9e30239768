"""Multi-GPU sharding of independent chains: one process per GPU, chains split into
contiguous ranges, no collective inside a sweep.  torch.distributed (NCCL over NVLink on the GPU
box, gloo in CPU tests) is used only to gather posterior summaries and particle-filter
log-normalisers.  Each rank passes chain_base = first global chain index to the engine, so the Philox
streams -- and therefore every sampled path -- are independent of how the chains are sharded."""
import numpy as np


def shard_range(total_chains, rank, world_size):
    """[lo, hi) of the chains owned by `rank`: contiguous, sizes differ by at most one"""
    base, extra = divmod(int(total_chains), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_work(n_per_chain, world_size):
    """contiguous ranges balanced by sum n^2 (the FFBS cost of a chain): boundaries [world+1]"""
    w = np.asarray(n_per_chain, dtype=np.float64) ** 2
    cum = np.concatenate([[0.0], np.cumsum(w)])
    targets = cum[-1] * np.arange(1, world_size) / world_size
    cuts = np.searchsorted(cum, targets, side="left")
    return np.concatenate([[0], cuts, [len(w)]]).astype(np.int64)


def _gather_rows(local, group=None):
    """all_gather of a [C_local, ...] tensor with possibly different C_local per rank -> [C_total, ...]"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)


def gather_summaries(time_in_state, jumps, group=None):
    """posterior summaries of all chains on every rank: (time_in_state [C_total,K] f64, jumps [C_total,K] i32)
    in global chain order (mcmc_utils.py:40-56 semantics per chain)."""
    return _gather_rows(time_in_state, group), _gather_rows(jumps, group)


def gather_lognorm(loglik, group=None):
    """particle-filter log-normalisers of all chains on every rank: [C_total]"""
    return _gather_rows(loglik.reshape(-1, 1), group).reshape(-1)


class SummaryGather:
    """All-gather of the per-chain posterior summaries ([C, 2K]: time in state, jump counts; mcmc_utils.py:40-56
    semantics per chain) of a sharded run, OVERLAPPED with the next sweep: `launch()` snapshots this rank's summaries
    into a staging buffer and enqueues one NCCL all-gather on a side stream, the sweep kernels keep the main stream;
    `result()` waits for the last one and returns them in global chain order.  Shard sizes follow shard_range(), so
    nothing is exchanged to learn them and nothing synchronises the host.  Without a process group (one GPU) the gather
    is the staging copy alone; on CPU tensors (gloo tests) everything is synchronous."""

    def __init__(self, chains_total, K, rank=0, world_size=1, device="cpu", group=None, overlap=True):
        import torch
        self.overlap = bool(overlap)  # False: the gather is ordered on the main stream, between two sweeps
        self.torch, self.K, self.rank, self.world, self.group = torch, int(K), int(rank), int(world_size), group
        self.sizes = [shard_range(chains_total, r, world_size)[1] - shard_range(chains_total, r, world_size)[0]
                      for r in range(world_size)]
        self.C = self.sizes[rank]
        self.m = max(self.sizes)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.stage = torch.zeros((self.m, 2 * self.K), dtype=torch.float64, device=self.device)
        self.out = torch.zeros((self.world * self.m, 2 * self.K), dtype=torch.float64, device=self.device)
        # high priority: its staging copies and the collective must not queue behind persistent sweep kernels
        self.stream = torch.cuda.Stream(self.device, priority=-1) if self.cuda else None
        self.work = None
        self.timing = []  # (start, end) CUDA events of every gather, on the side stream
        self.launched = 0
        self.bytes_per_gather = self.world * self.m * 2 * self.K * 8 if self.world > 1 else 0

    def reset_counters(self):
        """forget the gathers launched so far (warm-up) in `launched` / gather_ms(); call after result()"""
        self.timing = []
        self.launched = 0

    def launch(self, time_in_state, jumps):
        """enqueue a gather of the current summaries; returns at once"""
        t = self.torch
        import torch.distributed as dist
        if not self.cuda:
            self.stage[: self.C, : self.K] = time_in_state
            self.stage[: self.C, self.K:] = jumps.to(t.float64)
            if self.world > 1:
                dist.all_gather_into_tensor(self.out, self.stage, group=self.group)
            else:
                self.out.copy_(self.stage)
            self.launched += 1
            return self
        main = t.cuda.current_stream(self.device)
        if not self.overlap:
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            self.stage[: self.C, : self.K].copy_(time_in_state)
            self.stage[: self.C, self.K:].copy_(jumps)
            e0.record(main)
            if self.world > 1:
                dist.all_gather_into_tensor(self.out, self.stage, group=self.group)
            else:
                self.out.copy_(self.stage)
            e1.record(main)
            self.timing.append((e0, e1))
            self.launched += 1
            return self
        ready = t.cuda.Event()
        ready.record(main)
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        staged = t.cuda.Event()
        with t.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            if self.work is not None:
                self.work.wait()  # the previous gather's output buffer is about to be overwritten
            self.stage[: self.C, : self.K].copy_(time_in_state)
            self.stage[: self.C, self.K:].copy_(jumps)
            staged.record(self.stream)
            e0.record(self.stream)
            if self.world > 1:
                self.work = dist.all_gather_into_tensor(self.out, self.stage, group=self.group, async_op=True)
                self.work.wait()  # (stream-level: orders e1 after the collective, does not block the host)
            else:
                self.out.copy_(self.stage)
            e1.record(self.stream)
        main.wait_event(staged)  # the next sweep may overwrite the summaries once they are staged (microseconds)
        self.timing.append((e0, e1))
        self.launched += 1
        return self

    def result(self):
        """(time_in_state [C_total, K] f64, jumps [C_total, K] i32) of the last launched gather"""
        t = self.torch
        if self.cuda:
            self.stream.synchronize()
            if not self.overlap:
                t.cuda.current_stream(self.device).synchronize()
        rows = t.cat([self.out[r * self.m: r * self.m + s] for r, s in enumerate(self.sizes)], dim=0)
        return rows[:, : self.K].contiguous(), rows[:, self.K:].to(t.int32).contiguous()

    def gather_ms(self):
        """average duration of a gather on the side stream (staging excluded), ms"""
        if not self.timing:
            return 0.0
        self.stream.synchronize()
        t = self.torch
        t.cuda.current_stream(self.device).synchronize()
        return sum(a.elapsed_time(b) for a, b in self.timing) / len(self.timing)
