"""Multi-GPU: chains are independent (the reference runs them as separate sample() calls, util.py:386-389), so the
chain batch is sharded by rank -- contiguous blocks, chain c lives on rank floor(c*G/C) -- and every chain's random
stream is keyed by its GLOBAL id (Philox ``chain_offset``) or travels with it (injected mode).  Results therefore do
not depend on the number of GPUs.  There is no collective inside the sampling loop; what is collected afterwards is
a choice: per-chain summaries (tiny) or, with ``gather_samples=True``, every rank's sample block through ONE
``all_gather_into_tensor`` (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).

One process per GPU, launched by torchrun; ``torch.distributed`` must be initialised by the caller.
"""
import torch
import torch.distributed as dist


def shard_bounds(num_chains, rank, world):
    """Contiguous block partition: rank r owns chains [lo, hi).  Sizes differ by at most one."""
    lo = (num_chains * rank) // world
    hi = (num_chains * (rank + 1)) // world
    return lo, hi


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_rows(local, num_chains, group=None):
    """Gather row-sharded tensors (shard r = rows shard_bounds(num_chains, r, world)) into the full (num_chains, ...)
    tensor on every rank with ONE all_gather_into_tensor.  Ragged shards are padded to the largest shard."""
    rank, world = _world()
    if world == 1:
        return local
    sizes = [shard_bounds(num_chains, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] != mx:
        pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
        pad[:local.shape[0]] = local
    pad = pad.contiguous()
    out = pad.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(hi - lo == mx for lo, hi in sizes):
        return out
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)])


def pooled_moments(moment_sum, moment_sumsq, count_per_chain, group=None):
    """Posterior mean and variance pooled over ALL chains of ALL ranks from the sample sink's per-chain running sums
    (``sample_chains(..., moments=True)``: ``moment_sum`` / ``moment_sumsq`` (C_local, D), ``moment_count`` states per
    chain) with ONE all-reduce of a (2D+1,) fp64 vector -- the multi-GPU consumer that needs no sample gather at all
    (O(D) bytes over NVLink instead of O(C*S*D)).  Returns (mean (D,), var (D,), n) as fp64, identical on every rank;
    var is the population variance of the pooled draws."""
    s = moment_sum.double().sum(0)
    sq = moment_sumsq.double().sum(0)
    n = torch.tensor([float(moment_sum.shape[0]) * float(count_per_chain)], dtype=torch.float64, device=s.device)
    buf = torch.cat([s, sq, n])
    rank, world = _world()
    if world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    Dd = s.numel()
    total = float(buf[-1])
    if total <= 0:
        raise RuntimeError('pooled_moments: no post-burn states were accumulated')
    mean = buf[:Dd] / total
    var = buf[Dd:2 * Dd] / total - mean * mean
    return mean, var.clamp_min(0.0), total


def sample_chains_sharded(log_prob_func, params_init, gather_samples=False, runner=None, **kwargs):
    """``sample_chains`` over all ranks.  ``params_init`` is the FULL (C, D) batch on every rank (it is tiny next to
    the samples); each rank advances its block of chains on its own GPU.

    Returns a dict on every rank: ``num_rejected`` (C,), ``step_size`` (C,), ``bounds`` (this rank's [lo, hi)),
    ``local`` (this rank's HMCResult) and, when ``gather_samples``, ``samples`` (C, S-burn, D) collected with one
    all-gather.  With the sample sink's ``moments=True`` the per-chain running sums are pooled over all ranks by one
    O(D) all-reduce: ``posterior_mean`` / ``posterior_var`` (D,) fp64, ``posterior_n`` -- no sample ever leaves its GPU.  Injected-stream arguments ``normals`` (S, C, D) / ``log_uniforms`` (S, C) are sliced per rank.
    ``runner`` replaces ``samplers.sample_chains`` (used by the CPU tests of this host logic).
    """
    from . import samplers
    rank, world = _world()
    C = params_init.shape[0]
    lo, hi = shard_bounds(C, rank, world)
    kw = dict(kwargs)
    if hi == lo:
        raise RuntimeError('sample_chains_sharded: rank %d of %d would own no chain (C=%d < world); use fewer ranks'
                           % (rank, world, C))
    for name in ('normals', 'log_uniforms', 'perms', 'uniforms'):          # every injected stream is (S, C, ...)
        if kw.get(name) is not None:
            if kw[name].shape[1] != C:
                raise RuntimeError('%s must be (S, C=%d, ...), got %s' % (name, C, tuple(kw[name].shape)))
            kw[name] = kw[name][:, lo:hi]
    kw['chain_offset'] = kw.get('chain_offset', 0) + lo
    run = runner if runner is not None else samplers.sample_chains
    local = run(log_prob_func, params_init[lo:hi], **kw)
    out = {'bounds': (lo, hi), 'local': local,
           'num_rejected': all_gather_rows(local.num_rejected, C),
           'step_size': all_gather_rows(local.step_size, C)}
    if gather_samples:
        blk = local.samples_padded
        if not blk.is_cuda and dist.is_initialized() and dist.get_backend() == 'nccl':
            raise RuntimeError('gather_samples with store_on_GPU=False: the samples live in pinned host memory, which '
                               'NCCL cannot gather -- keep them on the GPU or gather on the host')
        out['samples'] = all_gather_rows(blk, C)[..., :local.dim]
    if getattr(local, 'moment_sum', None) is not None:          # sink moments requested: pool them over all ranks
        out['posterior_mean'], out['posterior_var'], out['posterior_n'] = pooled_moments(
            local.moment_sum, local.moment_sumsq, local.moment_count)
    return out
