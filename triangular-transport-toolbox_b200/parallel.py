"""
Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on the B200 box, gloo in CPU tests).

The reference's only parallel strategy is component parallelism over k with a process pool
(`Pool.starmap(worker_task, zip(np.flip(K), ...))`, tm.py:2789-2845).  Here the components are sharded
over the ranks, longest first, every rank fits its own components against its resident copy of the
ensemble, and ONE all-gather of a flat coefficient buffer follows (no per-iteration collective).
`inverse_map`/`map` shard by samples and need no collective at all.
"""

import numpy as np


def world():
    """(rank, world_size) of the default process group, (0, 1) if torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
    except Exception:
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_components(K, rank, size):
    """Components of `K` owned by `rank`: sorted by decreasing k (cost grows with k, the reference flips K
    for the same reason, tm.py:2814-2821) and dealt in snake order so the per-rank sums stay balanced."""
    order = sorted(K, reverse=True)
    mine = []
    for pos, k in enumerate(order):
        rnd, slot = divmod(pos, size)
        owner = slot if rnd % 2 == 0 else size - 1 - slot
        if owner == rank:
            mine.append(k)
    return mine


def owner_of(K, size):
    own = {}
    for r in range(size):
        for k in shard_components(K, r, size):
            own[k] = r
    return own


def allgather_coeffs(results, K, m_non, m_mon, device=None):
    """results: {k: (coeffs_nonmon, coeffs_mon)} for the locally fitted components.  Returns the same dict
    for ALL components of K after one all-gather of a flat [sum_k m_k] buffer per rank."""
    import torch
    import torch.distributed as dist
    rank, size = dist.get_rank(), dist.get_world_size()
    offs = np.concatenate(([0], np.cumsum([a + b for a, b in zip(m_non, m_mon)])))
    flat = np.zeros(int(offs[-1]))
    for i, k in enumerate(K):
        if k in results:
            flat[offs[i]:offs[i + 1]] = np.concatenate((results[k][0], results[k][1]))
    backend = dist.get_backend()
    dev = device if (backend == 'nccl' and device is not None) else torch.device('cpu')
    local = torch.from_numpy(flat).to(dev)
    gathered = [torch.empty_like(local) for _ in range(size)]
    dist.all_gather(gathered, local)
    own = owner_of(K, size)
    out = {}
    for i, k in enumerate(K):
        row = gathered[own[k]][offs[i]:offs[i + 1]].cpu().numpy()
        out[k] = (row[:m_non[i]].copy(), row[m_non[i]:].copy())
    return out


def allreduce_sum(vec, device=None):
    """Sample-sharded objective/gradient: sum a small float64 vector over the ranks (K < #GPUs case)."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = device if (backend == 'nccl' and device is not None) else torch.device('cpu')
    t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def warm_up(device=None):
    """Create the communicator and the all-reduce / all-gather channels now: the first NCCL collective of each
    kind costs seconds, which would otherwise be charged to the first optimize()."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = device if (backend == 'nccl' and device is not None) else torch.device('cpu')
    t = torch.zeros(8, dtype=torch.float64, device=dev)
    dist.all_reduce(t)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    if dev.type == 'cuda':
        torch.cuda.synchronize(dev)
