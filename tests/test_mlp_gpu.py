"""GPU parity tests of the Bayesian-NN path (define_model_log_prob, splitting integrators, sample_model /
sample_split_model, predict_model) against golden fixtures from the unmodified reference and the live oracle.

Tolerances: the likelihood gradient is a chain of GEMM-shaped sums whose order in the reference (CPU sgemm) is
unknowable, so states are compared to MLP_RTOL instead of bit-exactly; accept decisions must still be identical
(tests/parity.py reports the margin otherwise)."""
import os

import numpy as np
import pytest
import torch
import torch.utils.data as tud

import hamiltorch_b200 as hb
from hamiltorch_b200 import engine, targets as T, _native as N
from oracle import cases, hmc_oracle as O
from oracle.gen_golden import build_mlp_case
from tests import parity

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MLP_RTOL = 2e-4           # CEILING; the golden chains pass within 8x their measured error (<= 2e-7 -> 1e-5 floor, tests/parity.py)
SCHEME = {None: N.SCHEME_PLAIN, 'SPLITTING': N.SCHEME_SPLIT_SYM, 'SPLITTING_RAND': N.SCHEME_SPLIT_RAND,
          'SPLITTING_KMID': N.SCHEME_SPLIT_KMID}


def _autograd(f, q):
    q = q.detach().requires_grad_()
    lp = f(q)
    return torch.autograd.grad(lp, q)[0], lp.detach()


@pytest.mark.parametrize('name', sorted(cases.mlp_cases()))
def test_gradient_and_log_prob_match_autograd(name):
    """hmcx_grad_log_prob (hand-written backward) == autograd through the reference's closure, per split and summed."""
    case = cases.mlp_cases()[name]
    model, x, y, descs, inv_mass, tau_t = build_mlp_case(case)
    d = np.load(os.path.join(GOLD, name + '.npz'))
    q = torch.from_numpy(d['samples_0'])[:5]
    splits = descs if isinstance(descs, list) else [descs]
    for m, f in enumerate(splits):
        g, lp = engine.grad_log_prob(descs, q, split=m if isinstance(descs, list) else -1)
        for c in range(q.shape[0]):
            gr, lr = _autograd(f, q[c])
            scale = gr.abs().max().item()
            assert (g[c].cpu() - gr).abs().max().item() <= 2e-5 * scale, (name, m, c)
            assert abs(float(lp[c]) - float(lr)) <= 2e-5 * (abs(float(lr)) + 1)
    if isinstance(descs, list):
        g, lp = engine.grad_log_prob(descs, q, split=-1, want_grad=False)
        for c in range(q.shape[0]):
            tot = sum(float(f(q[c])) for f in descs)
            assert abs(float(lp[c]) - tot) <= 2e-5 * (abs(tot) + 1)


def test_gradient_at_config4_size():
    """BASELINE config 4 shapes: Linear(64,128)-ReLU-Linear(128,1), D=8449, N=1024 in 4 splits of 256."""
    model, x, y = cases.mlp_problem(seed=7, n=1024, n_in=64, hidden=128)
    descs = [T.MLPRegression.from_model(model, x[m * 256:(m + 1) * 256], y[m * 256:(m + 1) * 256], None, 100.,
                                        prior_scale=4) for m in range(4)]
    assert descs[0].dim == 8449
    torch.manual_seed(0)
    q = hb.util.flatten(model).detach()[None] + 0.05 * torch.randn(3, 8449)
    for m in (0, 3):
        g, lp = engine.grad_log_prob(descs, q, split=m)
        for c in range(3):
            gr, lr = _autograd(descs[m], q[c])
            assert (g[c].cpu() - gr).abs().max().item() <= 5e-5 * gr.abs().max().item()
            assert abs(float(lp[c]) - float(lr)) <= 5e-5 * (abs(float(lr)) + 1)


@pytest.mark.parametrize('name', sorted(cases.mlp_cases()))
def test_golden_chain_parity(name):
    """sample_model / sample_split_model chains of the reference, reproduced from the reference's random stream."""
    case = cases.mlp_cases()[name]
    model, x, y, descs, inv_mass, tau_t = build_mlp_case(case)
    d = np.load(os.path.join(GOLD, name + '.npz'))
    nC = len(case['seeds'])
    S, L, burn = case['num_samples'], case['num_steps_per_sample'], case['burn']
    init = torch.stack([torch.from_numpy(d['init_%d' % c]) for c in range(nC)])
    z = torch.stack([torch.from_numpy(d['z_%d' % c]) for c in range(nC)], 1)
    logu = torch.stack([torch.from_numpy(d['logu_%d' % c]) for c in range(nC)], 1)
    perms = torch.stack([torch.from_numpy(d['perms_%d' % c]) for c in range(nC)], 1)
    res = engine.hmc_run(descs, init, S, L, case['step_size'], burn=burn, inv_mass=inv_mass, normals=z,
                         log_uniforms=logu, perms=perms if case['scheme'] == 'SPLITTING_RAND' else None,
                         record_ham=True, scheme=SCHEME[case['scheme']])
    torch.cuda.synchronize()
    assert int(res.diverged.sum()) == 0
    for c in range(nC):
        parity.assert_chain_parity(res.samples[c].cpu().numpy(), res.accepted[c].cpu().numpy(),
                                   res.ham[c].cpu().numpy(), d['samples_%d' % c], d['accepted_%d' % c],
                                   d['ham_old_%d' % c], d['ham_new_%d' % c], d['logu_%d' % c], burn, exact=False,
                                   rtol=MLP_RTOL, tag='mlp/%s/c%d' % (name, c))


def test_sample_split_model_dropin_and_predict_model():
    """The reference-facing calls: hb.sample_split_model consumes torch's global stream like the reference (including
    the DataLoader's base-seed draw), hb.predict_model returns the reference's (S,N,O) tensor and log-prob list."""
    name = 'mlp_split_sym'
    case = cases.mlp_cases()[name]
    model, x, y, descs, inv_mass, tau_t = build_mlp_case(case)
    d = np.load(os.path.join(GOLD, name + '.npz'))
    seed = case['seeds'][0]
    D = descs[0].dim
    torch.manual_seed(seed)
    init = hb.util.flatten(model).detach().clone() + 0.05 * torch.randn(D)
    assert np.array_equal(init.numpy(), d['init_0'])
    M = case['num_splits']
    loader = tud.DataLoader(tud.TensorDataset(x, y), batch_size=x.shape[0] // M, shuffle=False)
    samples = hb.sample_split_model(model, loader, params_init=init, num_splits=M, model_loss='regression',
                                    num_samples=case['num_samples'], num_steps_per_sample=case['num_steps_per_sample'],
                                    step_size=case['step_size'], burn=case['burn'], tau_out=case['tau_out'],
                                    integrator=hb.Integrator.SPLITTING, verbose=False)
    got = torch.stack(samples).numpy()
    np.testing.assert_allclose(got, d['samples_0'], rtol=MLP_RTOL, atol=MLP_RTOL)
    pred, lps = hb.predict_model(model, [torch.from_numpy(s) for s in d['samples_0']], x=x, y=y,
                                 model_loss='regression', tau_out=case['tau_out'])
    assert pred.shape == d['pred'].shape and len(lps) == d['pred'].shape[0] and lps[0].shape == (1,)
    np.testing.assert_allclose(pred.numpy(), d['pred'], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(torch.stack(lps).numpy(), d['pred_log_prob'], rtol=2e-5)
    with pytest.raises(NotImplementedError):
        hb.sample_model(model, x, y, init, model_loss=lambda out, tgt: ((out - tgt) ** 2).sum(1))   # callable loss
    with pytest.raises(RuntimeError, match='greater than length 1'):
        hb.sample(descs[:1], init, integrator=hb.Integrator.SPLITTING, rng='philox')


def test_config4_shapes_philox_properties():
    """BASELINE config 4 shapes on one GPU, 8 chains (the per-GPU share of the 64-chain run), short: symmetric
    split HMC with the survey's settings accepts ~0.97; rejected stored iterations repeat the previous row."""
    model, x, y = cases.mlp_problem(seed=0, n=1024, n_in=64, hidden=128)
    M = 4
    descs = [T.MLPRegression.from_model(model, x[m * 256:(m + 1) * 256], y[m * 256:(m + 1) * 256], None, 100.,
                                        prior_scale=M) for m in range(M)]
    D = descs[0].dim
    init = hb.util.flatten(model).detach()[None] + 0.01 * torch.randn(8, D, generator=torch.Generator().manual_seed(0))
    res = hb.sample_chains(descs, init, num_samples=12, num_steps_per_sample=10, step_size=5e-4,
                           inv_mass=torch.ones(D), integrator=hb.Integrator.SPLITTING, rng='philox', seed=3,
                           record_ham=True)
    torch.cuda.synchronize()
    acc = res.accepted.cpu().bool()
    assert int(res.diverged.sum()) == 0
    assert acc.float().mean().item() > 0.7
    s = res.samples.cpu()
    assert torch.equal(s[:, 0], init)
    same = (s[:, 2:] == s[:, 1:-1]).all(-1)
    assert torch.equal(same, ~acc[:, 2:])
    dH = (res.ham[..., 1] - res.ham[..., 0]).cpu()
    assert torch.isfinite(dH).all() and dH.abs().median() < 2.0


@pytest.mark.parametrize('name', ['mlp_binary', 'mlp_multiclass', 'mlp_logsoftmax_split'])
def test_classification_predict_model(name):
    """predict_model for the classification likelihoods: logits (or log-probs for a LogSoftmax model) and the 0-d
    log-prob per sample, vs the reference's own predict_model output."""
    case = cases.mlp_cases()[name]
    model, x, y, descs, inv_mass, tau_t = build_mlp_case(case)
    d = np.load(os.path.join(GOLD, name + '.npz'))
    pred, lps = hb.predict_model(model, [torch.from_numpy(s) for s in d['samples_0']], x=x, y=y,
                                 model_loss=case['model_loss'], tau_out=case['tau_out'])
    assert pred.shape == d['pred'].shape and lps[0].shape == ()
    np.testing.assert_allclose(pred.numpy(), d['pred'], rtol=5e-5, atol=5e-5)
    np.testing.assert_allclose(torch.stack(lps).numpy(), d['pred_log_prob'].reshape(-1), rtol=5e-5)


@pytest.mark.parametrize('rows_per_split,expect_ctas', [(192, 2), (320, 4)])
def test_cluster_split_chain_parity_vs_live_oracle(rows_per_split, expect_ctas):
    """Splits with several 64-row tiles are shared by a thread-block cluster (2 or 4 CTAs per chain, partial gradients
    and log-likelihood sums combined through distributed shared memory in rank order): same chains as the oracle."""
    M, S, L, burn = 2, 8, 3, 1
    model, x, y = cases.mlp_problem(seed=11, n=M * rows_per_split, n_in=6, hidden=16)
    descs = [T.MLPTarget.from_model(model, x[m * rows_per_split:(m + 1) * rows_per_split],
                                    y[m * rows_per_split:(m + 1) * rows_per_split], None, 20., prior_scale=M)
             for m in range(M)]
    D = descs[0].dim
    C = 3
    inits, zs, lus = [], [], []
    for seed in range(C):
        init, z, logu, _ = O.reference_stream(40 + seed, D, S,
                                              prior=lambda: hb.util.flatten(model).detach() + 0.05 * torch.randn(D))
        inits.append(init), zs.append(z), lus.append(logu)
    res = engine.hmc_run(descs, torch.stack(inits), S, L, 0.004, burn=burn, normals=torch.stack(zs, 1),
                         log_uniforms=torch.stack(lus, 1), record_ham=True, scheme=N.SCHEME_SPLIT_SYM)
    torch.cuda.synchronize()
    assert int(res.diverged.sum()) == 0
    for c in range(C):
        o = O.sample_hmc(descs, inits[c], num_samples=S, num_steps_per_sample=L, step_size=0.004, burn=burn,
                         split_scheme=O.SPLIT_SYM, normals=zs[c], log_uniforms=lus[c])
        parity.assert_chain_parity(res.samples[c].cpu().numpy(), res.accepted[c].cpu().numpy(),
                                   res.ham[c].cpu().numpy(), torch.stack(o['samples']).numpy(), o['accepted'],
                                   o['ham_old'], o['ham_new'], lus[c].numpy(), burn, exact=False, rtol=MLP_RTOL)


def test_pinned_cluster_size_is_bit_reproducible_across_chain_counts():
    """hmcx_mlp_t.cluster_size pins the CTAs per chain, so a chain's bits do not depend on how many chains share the
    launch (the automatic choice shrinks the cluster when chains outnumber SMs)."""
    M, S, L, rows = 2, 6, 3, 320
    model, x, y = cases.mlp_problem(seed=12, n=M * rows, n_in=6, hidden=16)
    descs = [T.MLPTarget.from_model(model, x[m * rows:(m + 1) * rows], y[m * rows:(m + 1) * rows], None, 20.,
                                    prior_scale=M) for m in range(M)]
    descs[0].cluster_size = 2
    D = descs[0].dim
    g = torch.Generator().manual_seed(3)
    big = 100
    init = hb.util.flatten(model).detach()[None] + 0.05 * torch.randn(big, D, generator=g)
    z = torch.randn(S, big, D, generator=g)
    lu = torch.log(torch.rand(S, big, generator=g))
    run = lambda c: engine.hmc_run(descs, init[:c], S, L, 0.004, normals=z[:, :c].contiguous(),
                                   log_uniforms=lu[:, :c].contiguous(), scheme=N.SCHEME_SPLIT_SYM)
    a, b = run(3), run(big)
    torch.cuda.synchronize()
    assert torch.equal(a.samples, b.samples[:3]) and torch.equal(a.accepted, b.accepted[:3])


# ----------------------------------------------------------------------------------------------------------
# stand-alone samplers.leapfrog with a SPLITTING integrator (samplers.py:494-603)
# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['mlp_split_sym', 'mlp_split_rand', 'mlp_split_kmid', 'mlp_deep_tanh_mass'])
def test_standalone_split_leapfrog_matches_the_reference(name):
    from oracle.gen_golden import build_mlp_case
    case = cases.mlp_cases()[name]
    d = np.load(os.path.join(GOLD, 'split_standalone.npz'))
    model, x, y, descs, inv_mass, tau_t = build_mlp_case(case)
    q, p = torch.from_numpy(d[name + '.q0']), torch.from_numpy(d[name + '.p0'])
    integ = getattr(hb.Integrator, case['scheme'])
    qs, ps = hb.leapfrog(q, p, descs, steps=3, step_size=case['step_size'], inv_mass=inv_mass, sampler=hb.Sampler.HMC,
                         integrator=integ, rng_perms=torch.from_numpy(d[name + '.perm'])[None])
    assert len(qs) == len(ps) == 3 and tuple(qs[0].shape) == (q.numel(),)
    scale_q, scale_p = np.abs(d[name + '.q_traj']).max(), np.abs(d[name + '.p_traj']).max()
    assert np.abs(torch.stack(qs).cpu().numpy() - d[name + '.q_traj']).max() <= MLP_RTOL * scale_q
    assert np.abs(torch.stack(ps).cpu().numpy() - d[name + '.p_traj']).max() <= MLP_RTOL * scale_p
    # batched: C chains at once, chain c == the single call
    qb, pb = q.repeat(3, 1), p.repeat(3, 1)
    qs3, ps3 = hb.leapfrog(qb, pb, descs, steps=3, step_size=case['step_size'], inv_mass=inv_mass, sampler=hb.Sampler.HMC,
                           integrator=integ, rng_perms=torch.from_numpy(d[name + '.perm'])[None].repeat(3, 1))
    assert tuple(qs3[0].shape) == (3, q.numel())
    assert torch.equal(qs3[-1][1].cpu(), qs[-1].cpu()) and torch.equal(ps3[-1][2].cpu(), ps[-1].cpu())


def test_standalone_split_leapfrog_argument_errors():
    d = cases.mlp_cases()['mlp_split_sym']
    from oracle.gen_golden import build_mlp_case
    model, x, y, descs, inv_mass, tau_t = build_mlp_case(d)
    q = hb.util.flatten(model).detach()
    with pytest.raises(RuntimeError):                             # :466-467
        hb.leapfrog(q, torch.zeros_like(q), descs[0], sampler=hb.Sampler.HMC, integrator=hb.Integrator.SPLITTING)
    with pytest.raises(RuntimeError):                             # :497-498
        hb.leapfrog(q, torch.zeros_like(q), descs[:1], sampler=hb.Sampler.HMC, integrator=hb.Integrator.SPLITTING)


def test_predict_model_loader_with_more_batches_than_one_launch_takes():
    """A DataLoader with 70 batches (> the 64 splits of one launch): launches of 64 + 6 batches, predictions concatenated in
    batch order, log-probs added -- against the oracle's restatement of the reference loop (samplers.py:1527-1540)."""
    import torch.utils.data as tud
    model, x, y = cases.mlp_problem(seed=3, n=140, n_in=5, hidden=16)
    loader = tud.DataLoader(tud.TensorDataset(x, y), batch_size=2, shuffle=False)
    D = hb.util.flatten(model).numel()
    g = torch.Generator().manual_seed(1)
    samples = [hb.util.flatten(model).detach() + 0.1 * torch.randn(D, generator=g) for _ in range(3)]
    pred, lps = hb.predict_model(model, [s.cuda() for s in samples], test_loader=loader, model_loss='regression', tau_out=10.)
    assert pred.shape == (3, 140, 1) and len(lps) == 3
    # the reference's sum over the 70 batch closures, each with the prior divided by the number of batches (:1527)
    for i, s in enumerate(samples):
        lp_ref = float(sum(T.MLPRegression.from_model(model, x[2 * b:2 * b + 2], y[2 * b:2 * b + 2], None, 10., prior_scale=70)(s)
                           for b in range(70)))
        assert abs(float(lps[i]) - lp_ref) <= 5e-5 * (1 + abs(lp_ref))
    whole, _ = hb.predict_model(model, [s.cuda() for s in samples], x=x.cuda(), y=y.cuda(), model_loss='regression', tau_out=10.)
    assert torch.allclose(pred.cpu(), whole.cpu(), rtol=1e-5, atol=1e-6)
