"""GPU: the sample sink of the persistent HMC kernel (include/hmcx.h hmcx_sink_t / hmcx_hmc_run_sink): thinning, running
moments in registers, no-sample runs and the store_on_GPU=False path (kernel streams retained rows into pinned host
memory).  Everything is checked bit-exactly against the plain run of the same chains (same Philox stream)."""
import pytest
import torch

import hamiltorch_b200 as hb
from hamiltorch_b200 import targets as T

pytestmark = pytest.mark.gpu

KW = dict(num_samples=50, num_steps_per_sample=4, step_size=0.25, burn=7, rng='philox', seed=21)


def _setup(D=300, C=6):
    g = torch.Generator().manual_seed(3)
    tgt = T.GaussianDiag(torch.randn(D, generator=g), 0.3 + torch.rand(D, generator=g))
    init = tgt.mean[None] + 0.3 * torch.randn(C, D, generator=g)
    im = 0.5 + torch.rand(D, generator=g)
    return tgt, init, im


@pytest.mark.parametrize('nuts', [False, True])
def test_thinning_and_moments_equal_the_plain_run(nuts):
    tgt, init, im = _setup()
    kw = dict(KW, inv_mass=im, sampler=hb.Sampler.HMC_NUTS if nuts else hb.Sampler.HMC)
    full = hb.sample_chains(tgt, init, **kw)
    thin = hb.sample_chains(tgt, init, thin=4, moments=True, **kw)
    none = hb.sample_chains(tgt, init, keep_samples=False, moments=True, **kw)
    torch.cuda.synchronize()
    assert torch.equal(thin.accepted, full.accepted) and torch.equal(thin.step_size, full.step_size)
    assert thin.samples.shape[1] == 1 + (50 - 7 - 1) // 4
    assert torch.equal(thin.samples, full.samples[:, ::4])
    # moments: compensated in-kernel sums (hi + lo) == fp64 sums over the reference's returned list minus element 0
    x = full.samples[:, 1:].double()
    s, sq = x.sum(1), (x * x).sum(1)
    assert thin.moment_count == full.samples.shape[1] - 1
    assert thin.moment_sum.dtype == torch.float64
    assert torch.allclose(thin.moment_sum, s, rtol=1e-10, atol=1e-10)
    assert torch.allclose(thin.moment_sumsq, sq, rtol=1e-10, atol=1e-10)
    assert torch.equal(none.moment_sum, thin.moment_sum) and torch.equal(none.moment_sumsq, thin.moment_sumsq)
    assert torch.equal(none.final_state, full.samples[:, -1])
    with pytest.raises(RuntimeError):
        none.samples


def test_store_on_gpu_false_streams_samples_to_pinned_host_memory():
    tgt, init, im = _setup(D=1024, C=16)
    full = hb.sample_chains(tgt, init, **KW)
    host = hb.sample_chains(tgt, init, store_on_GPU=False, **KW)
    torch.cuda.synchronize()
    assert not host.samples.is_cuda and host.samples_padded.is_pinned()
    assert torch.equal(host.samples, full.samples.cpu())
    # the reference-shaped entry point: list of CPU tensors (samplers.py:1008-1012)
    torch.manual_seed(0)
    a = hb.sample(tgt, init[0], num_samples=12, num_steps_per_sample=3, step_size=0.2, store_on_GPU=False, verbose=False)
    assert len(a) == 12 and all(not t.is_cuda for t in a)


def test_sink_is_refused_where_it_is_not_implemented():
    D = 24
    cov = torch.eye(D, dtype=torch.float64) * 2
    with pytest.raises(NotImplementedError):
        hb.sample_chains(T.GaussianFull(torch.zeros(D), cov=cov), torch.zeros(2, D), num_samples=5, thin=2)


def test_sink_edge_cases():
    """burn = S-1 (nothing but params_init is ever retained: zero moment count), thin larger than the run, D % 4 != 0."""
    g = torch.Generator().manual_seed(9)
    D, C = 37, 3
    tgt = T.GaussianDiag(torch.zeros(D), 0.5 + torch.rand(D, generator=g))
    init = 0.2 * torch.randn(C, D, generator=g)
    r = hb.sample_chains(tgt, init, num_samples=6, num_steps_per_sample=3, step_size=0.2, burn=5, moments=True, seed=2)
    torch.cuda.synchronize()
    assert r.samples.shape == (C, 1, D) and torch.equal(r.samples[:, 0].cpu(), init)
    assert r.moment_count == 0 and float(r.moment_sum.abs().sum()) == 0.0
    full = hb.sample_chains(tgt, init, num_samples=12, num_steps_per_sample=3, step_size=0.2, burn=2, seed=2)
    thin = hb.sample_chains(tgt, init, num_samples=12, num_steps_per_sample=3, step_size=0.2, burn=2, seed=2, thin=50,
                            moments=True)
    torch.cuda.synchronize()
    assert thin.samples.shape == (C, 1, D) and torch.equal(thin.samples[:, 0], full.samples[:, 0])
    assert torch.allclose(thin.moment_sum, full.samples[:, 1:].double().sum(1), rtol=1e-12, atol=1e-12)
    with pytest.raises(RuntimeError):
        hb.sample_chains(tgt, init, num_samples=5, thin=0)


def test_moments_of_a_long_run_far_from_zero():
    """|mean| = 100 std over 20000 post-burn iterations.  A naive fp32 running sum of x^2 loses
    percents of the variance here (error ~ n*eps); the compensated sums carry ~ n*eps^2 (measured 1.6e-7 on the variance at
    n = 2e4, mean^2/var = 1e4), i.e. the posterior variance of every chain to 2e-6 of what fp64 arithmetic over the stored
    samples gives -- and pooled_moments (var = E[x^2] - mean^2 in fp64) must agree."""
    from hamiltorch_b200 import distributed
    D, C, S = 16, 4, 20001
    tgt = T.GaussianDiag(torch.full((D,), 100.0), torch.ones(D))
    init = tgt.mean[None] + torch.randn(C, D, generator=torch.Generator().manual_seed(1))
    kw = dict(num_samples=S, num_steps_per_sample=3, step_size=0.4, rng='philox', seed=5)
    full = hb.sample_chains(tgt, init, **kw)
    mom = hb.sample_chains(tgt, init, keep_samples=False, moments=True, **kw)
    torch.cuda.synchronize()
    x = full.samples[:, 1:].double()
    n = x.shape[1]
    assert mom.moment_count == n
    mean64, var64 = x.mean(1), x.var(1, unbiased=False)
    mean = mom.moment_sum / n
    var = mom.moment_sumsq / n - mean * mean
    assert torch.allclose(mean, mean64, rtol=1e-9)
    assert torch.allclose(var, var64, rtol=2e-6), (var - var64).abs().max()
    assert 0.8 < float(var64.mean()) < 1.25                   # the chain did sample N(100, 1)
    # the naive fp32 recurrence the kernel used in round 1, for the record: off by > 1e-3 here
    s32 = torch.zeros_like(full.samples[:, 0]); q32 = torch.zeros_like(s32)
    for j in range(1, 2001):
        s32 = s32 + full.samples[:, j]; q32 = q32 + full.samples[:, j] * full.samples[:, j]
    naive = q32.double() / 2000 - (s32.double() / 2000) ** 2
    exact = full.samples[:, 1:2001].double().var(1, unbiased=False)
    assert float((naive - exact).abs().max()) > 10 * float((var - var64).abs().max())
    pm, pv, pn = distributed.pooled_moments(mom.moment_sum, mom.moment_sumsq, mom.moment_count)
    allx = x.reshape(-1, D)
    assert pn == C * n and torch.allclose(pm, allx.mean(0), rtol=1e-9)
    assert torch.allclose(pv, allx.var(0, unbiased=False), rtol=2e-6)


@pytest.mark.parametrize('nuts,windows,burn', [(False, 4, 7), (True, 3, 20), (False, 50, 0), (False, 7, 48)])
def test_windowed_delivery_into_a_pinned_host_block_equals_the_plain_run(nuts, windows, burn):
    """out=<pinned host block>, host_windows=W: the run is cut into W windows of iterations (hmcx_hmc_run with iter_begin /
    iter_end) and each window's sample slots are delivered by hmcx_copy_rows_async on a second stream while the next window
    computes: the same bytes as one launch, the same flags / step sizes (windows chain through q_cur, eps, the NUTS state)."""
    tgt, init, im = _setup()
    kw = dict(KW, inv_mass=im, burn=burn, sampler=hb.Sampler.HMC_NUTS if nuts else hb.Sampler.HMC, record_ham=True)
    if nuts and burn < 1:
        pytest.skip('NUTS needs burn >= 1')
    full = hb.sample_chains(tgt, init, **kw)
    S, C = KW['num_samples'], init.shape[0]
    ld = full.samples_padded.shape[-1]
    host = torch.full((C, S - burn, ld), float('nan')).pin_memory()
    win = hb.sample_chains(tgt, init, out=host, host_windows=windows, **kw)
    torch.cuda.synchronize()
    assert not win.samples_padded.is_cuda and win.samples_padded.data_ptr() == host.data_ptr()
    assert torch.equal(host, full.samples_padded.cpu())
    assert torch.equal(win.accepted, full.accepted) and torch.equal(win.ham, full.ham)
    assert torch.equal(win.step_size, full.step_size) and torch.equal(win.num_rejected, full.num_rejected)
