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
