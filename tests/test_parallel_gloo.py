"""Multi-process plumbing on the CPU (gloo, world_size 2): component sharding and the single all-gather of
coefficients that follows the fits (the N>1 path of optimize()); sample-sharded all-reduce helper."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import transport_map  # noqa: F401  (registers the package as ttt_b200)
from ttt_b200.parallel import shard_components, owner_of, allgather_coeffs, allreduce_sum


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_sharding_is_a_balanced_partition():
    K = list(range(64))
    for size in (1, 2, 3, 4, 8):
        shards = [shard_components(K, r, size) for r in range(size)]
        assert sorted(k for s in shards for k in s) == K
        loads = [sum(100 + k for k in s) for s in shards]           # cost grows with k
        assert max(loads) - min(loads) <= 100 + 63
        assert all(len(s) in (64 // size, 64 // size + 1) for s in shards)
        own = owner_of(K, size)
        assert all(k in shards[own[k]] for k in K)
    assert shard_components([5, 2, 9], 0, 1) == [9, 5, 2]            # longest first (tm.py:2821)


def _worker(rank, size, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=size)
    K = list(range(7))
    m_non = [1 + 3 * k for k in K]
    m_mon = [3 if k == 0 else 4 for k in K]
    mine = shard_components(K, rank, size)
    fake = lambda k: (np.arange(m_non[k]) + 100.0 * k, -np.arange(m_mon[k]) - 10.0 * k)
    gathered = allgather_coeffs({k: fake(k) for k in mine}, K, m_non, m_mon)
    ok = all(np.array_equal(gathered[k][0], fake(k)[0]) and np.array_equal(gathered[k][1], fake(k)[1]) for k in K)
    total = allreduce_sum(np.array([1.0 + rank, 2.0 * rank]))
    ok = ok and np.allclose(total, [sum(1.0 + r for r in range(size)), sum(2.0 * r for r in range(size))])
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_allgather_of_coefficients_world_size_2():
    size, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(size, port, ret), nprocs=size, join=True)
        assert all(ret.get(r) for r in range(size))


def _lockstep_worker(rank, size, port, ret):
    """Sample-sharded separable fit, host side: every rank holds a shard of the samples, each evaluation all-reduces
    the partial sums (sum log dS, sum dPsi / dS) and the ranks advance the SAME L-BFGS-B iteration in lockstep
    (transport_map._fit_separable_lockstep with sample_sharded=True; the sums come from K-sepobj there, from numpy
    here).  Every rank must end on the single-process fit."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=size)
    from ttt_b200 import hostopt
    rng = np.random.default_rng(11)
    n, comps = 600, 3
    probs = []
    for c in range(comps):
        m = 2 + c
        dPsi = np.abs(rng.standard_normal((n, m))) + 0.1            # derivative basis (positive)
        A = rng.standard_normal((m, m))
        A = A @ A.T / m + 0.2 * np.eye(m)
        probs.append((dPsi, A, 1e-8 * A.sum(axis=1), rng.uniform(0.2, 1.0, m)))
    lo, hi = rank * n // size, (rank + 1) * n // size

    def sums(c, b, rows):
        dPsi = probs[c][0][rows]
        dS = dPsi @ (b + 1e-8)
        return np.concatenate(([np.sum(np.log(dS))], (dPsi / dS[:, None]).sum(axis=0)))

    def assemble(c, b, out):
        _, A, bvec, _ = probs[c]
        Ax = A @ b
        return b @ Ax / 2 - out[0] / n + b @ bvec, Ax - out[1:] / n + bvec

    held = {}

    def launch(i, b):
        held[i] = np.array(b)

    def collect(i):
        return assemble(i, held[i], allreduce_sum(sums(i, held[i], slice(lo, hi))))

    bounds = [(np.zeros(len(p[3])), np.full(len(p[3]), np.inf)) for p in probs]
    got = hostopt.lbfgsb_lockstep([p[3] for p in probs], bounds, launch, collect)
    whole = hostopt.lbfgsb_lockstep([p[3] for p in probs], bounds, lambda i, b: held.__setitem__(i, np.array(b)),
                                    lambda i: assemble(i, held[i], sums(i, held[i], slice(0, n))))
    ok = all(np.allclose(g.x, w.x, rtol=1e-9, atol=1e-12) and g.status == 0 for g, w in zip(got, whole))
    # identical iterates on every rank (the all-reduced sums are the same numbers everywhere)
    flat = np.concatenate([g.x for g in got])
    both = [torch.zeros(flat.size, dtype=torch.float64) for _ in range(size)]
    dist.all_gather(both, torch.from_numpy(flat))
    ok = ok and all(torch.equal(both[0], t) for t in both)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sample_sharded_lockstep_fit_world_size_2():
    size, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_lockstep_worker, args=(size, port, ret), nprocs=size, join=True)
        assert all(ret.get(r) for r in range(size))
