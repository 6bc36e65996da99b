"""CPU, world_size 2, gloo: the host-side multi-GPU logic (chain sharding by rank, global chain ids, the single
all-gather, ragged shards).  The kernels themselves cannot run here; a stand-in runner produces, for every chain,
values that depend only on its GLOBAL id -- exactly the invariant the Philox keying gives the real kernels (pinned
on the GPU by tests/test_hmc_gpu.py::test_philox_is_reproducible_and_sharding_invariant)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hamiltorch_b200 import distributed as D


class FakeResult:
    def __init__(self, samples, rej, eps):
        self.samples_padded, self.num_rejected, self.step_size, self.dim = samples, rej, eps, samples.shape[-1] - 1


def fake_runner(log_prob_func, q0, num_samples=4, chain_offset=0, **kw):
    Cl, Dd = q0.shape
    ids = torch.arange(chain_offset, chain_offset + Cl, dtype=torch.float32)
    samples = torch.zeros(Cl, num_samples, Dd + 1)                      # +1: a pad column, like ld > D
    samples[..., :Dd] = q0[:, None, :] + ids[:, None, None] * 1000 + torch.arange(num_samples)[None, :, None]
    return FakeResult(samples, (ids * 3).int(), ids / 7)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, C, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        q0 = torch.arange(C * 5, dtype=torch.float32).reshape(C, 5)
        out = D.sample_chains_sharded(None, q0, gather_samples=True, runner=fake_runner, num_samples=4)
        ref = fake_runner(None, q0, num_samples=4)
        ok = torch.equal(out['samples'], ref.samples_padded[..., :5]) \
            and torch.equal(out['num_rejected'], ref.num_rejected) \
            and torch.equal(out['step_size'], ref.step_size) \
            and out['bounds'] == D.shard_bounds(C, rank, world)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('C', [8, 7])          # even and ragged shards
def test_sharded_equals_single_process_gloo_world2(C):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, C, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_bounds_cover_all_chains():
    for C in (1, 7, 8, 256, 1024):
        for world in (1, 2, 3, 8):
            b = [D.shard_bounds(C, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == C
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def _moments_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        draws = torch.randn(6, 9, 4, generator=g) * 2 + 1          # 6 chains x 9 post-burn states x D=4, same on all ranks
        lo, hi = D.shard_bounds(6, rank, world)
        mine = draws[lo:hi]
        mean, var, n = D.pooled_moments(mine.sum(1), (mine * mine).sum(1), 9)
        flat = draws.reshape(-1, 4).double()
        ok = n == 54 and torch.allclose(mean, flat.mean(0), atol=1e-6) and \
            torch.allclose(var, flat.var(0, unbiased=False), atol=1e-5)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_pooled_moments_all_reduce_gloo_world2():
    """The sink's per-chain running sums pooled over ranks by one O(D) all-reduce == moments of all draws."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_moments_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    got = dict(q.get(timeout=10) for _ in range(2))
    assert got == {0: True, 1: True}
