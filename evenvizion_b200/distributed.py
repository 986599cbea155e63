"""Sharding of a frame chain over the GPUs of one box (one process per GPU).

Pairs are independent units (match, RANSAC, refit, static filter), so each rank takes a
contiguous range of pairs and needs the frames [start, end] inclusive (a one-frame halo that the
host loader supplies).  The only exchange is the cumulative superposition: every rank reduces
its shard to a 160-byte summary (evz_chain_scan, include/evz.h), the summaries are all-gathered
(NCCL over NVLink for CUDA tensors, gloo in the CPU tests), and each rank derives the seeds that
make its local scan a slice of the global one.
"""
import numpy as np


def shard_range(n_pairs, rank, world):
    """Contiguous pair range [lo, hi) of `rank`."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _norm(m):
    return m / m[2, 2]


def seeds_from_summaries(summaries, rank, policy=True):
    """summaries: list over ranks of the 20-double shard summary.  Returns (seed_S (3,3),
    seed_G (3,3) or None) for `rank`: the superposition before its first pair and the last valid
    step matrix before it (None-H forward fill across the shard boundary)."""
    S = np.eye(3)
    last_G = None
    for r in range(rank):
        s = np.asarray(summaries[r], np.float64)
        R, G_last, lead, has = s[0:9].reshape(3, 3), s[9:18].reshape(3, 3), int(round(s[18])), s[19] > 0.5
        if policy and last_G is not None:
            for _ in range(lead):                    # leading failures of shard r repeat the carried step
                S = _norm(S @ last_G)
        if has:
            S = _norm(S @ R)
            last_G = G_last
    return S, (last_G if policy else None)


def all_gather_summaries(summary, group=None):
    """All-gather the 20-double shard summary with torch.distributed (NCCL for CUDA tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bufs = [torch.empty_like(summary) for _ in range(world)]
    dist.all_gather(bufs, summary, group=group)
    return [b.cpu().numpy() for b in bufs]


def scan_sharded(eng, G, status, policy=True, group=None, want_fixed=True):
    """Cumulative superposition of a video whose pairs are sharded over the ranks of `group`, without leaving the
    device: unseeded local scan + 20-double summary (evz_chain_scan), ONE all-gather of the summaries
    (`all_gather_into_tensor`: NCCL over NVLink), then evz_chain_seed_apply.  Returns (S, H_fixed) of this rank's pairs."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        S, Hf, _ = eng.chain_scan(G, status, policy, want_fixed=want_fixed)
        return S, Hf
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    S, _, summ = eng.chain_scan(G, status, policy, want_fixed=False, want_summary=True)
    buf = torch.empty((world, 20), dtype=torch.float64, device=summ.device)
    dist.all_gather_into_tensor(buf, summ, group=group)
    S, Hf, _ = eng.chain_seed_apply(buf, rank, S, policy, want_fixed)
    return S, Hf
