"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden

For every case in oracle/cases.py and every chain seed it
  1. runs ``hamiltorch.sample`` (the reference) with torch seeded as util.setup_chain does (util.py:386-389),
  2. replays the same seed through the oracle and ASSERTS the oracle's chain is bit-identical to the
     reference's (this is what pins the oracle),
  3. stores: the reference samples, the pre-drawn random stream the reference consumed (standard normals and
     log-uniforms), and the per-iteration accept flags / Hamiltonians / step sizes, which the
     reference computes but does not return (taken from the bit-identical oracle replay).
It also stores the reference's own hot-path unit test (tests/test_util.py:97-110) as a trajectory fixture.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import cases, hmc_oracle as O          # noqa: E402
from oracle import rmhmc_oracle as R               # noqa: E402
from oracle.ref_import import import_reference     # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def run_plain_case(ref, name, case):
    tgt, kw, dim = case['target'], dict(case['kw']), case['target'].dim
    nuts = kw.pop('nuts', False)
    S = kw['num_samples']
    out = {}
    for ci, seed in enumerate(case['seeds']):
        init = cases.make_init(case['init'], dim, seed)
        ref_kw = dict(kw)
        if nuts:
            ref_kw['sampler'] = ref.Sampler.HMC_NUTS
        samples, extra = ref.sample(log_prob_func=tgt, params_init=init, debug=2, verbose=False, **ref_kw)
        samples = torch.stack(samples)
        # oracle replay from the same RNG state
        init2 = cases.make_init(case['init'], dim, seed)
        assert torch.equal(init, init2)
        okw = dict(kw)
        res = O.sample_hmc(tgt, init2, nuts=nuts, **okw)
        osamples = torch.stack(res['samples'])
        assert osamples.shape == samples.shape, (name, osamples.shape, samples.shape)
        assert torch.equal(osamples, samples), '%s seed %d: oracle != reference (max abs %g)' % (
            name, seed, (osamples - samples).abs().max())
        if nuts:
            assert extra == res['step_size'], (extra, res['step_size'])
        else:
            assert abs(extra - (1 - res['num_rejected'] / S)) < 1e-12
        # the stream the reference consumed
        torch.manual_seed(seed)
        cases.make_init(case['init'], dim, seed)
        z = torch.empty(S, dim)
        logu = torch.empty(S)
        blocks = [b.shape[0] for b in kw['inv_mass']] if isinstance(kw.get('inv_mass'), list) else None
        for n in range(S):
            z[n] = torch.cat([torch.randn(b) for b in blocks]) if blocks else torch.randn(dim)
            if res['diverged'][n]:
                logu[n] = 0.0          # a LogProbError iteration never reaches torch.rand(1) (samplers.py:1004 vs :1045)
            else:
                logu[n] = torch.log(torch.rand(1))[0]
        # and check that the oracle driven by the injected stream is again identical
        res2 = O.sample_hmc(tgt, init2, nuts=nuts, normals=z, log_uniforms=logu, **okw)
        assert torch.equal(torch.stack(res2['samples']), samples)
        out['init_%d' % ci] = init.numpy()
        out['samples_%d' % ci] = samples.numpy()
        out['z_%d' % ci] = z.numpy()
        out['logu_%d' % ci] = logu.numpy()
        out['accepted_%d' % ci] = np.array(res['accepted'], dtype=np.uint8)
        out['diverged_%d' % ci] = np.array(res['diverged'], dtype=np.uint8)
        out['ham_old_%d' % ci] = np.array(res['ham_old'], dtype=np.float64)
        out['ham_new_%d' % ci] = np.array(res['ham_new'], dtype=np.float64)
        out['step_sizes_%d' % ci] = np.array(res['step_sizes'], dtype=np.float64)
        out['final_step_size_%d' % ci] = np.float64(res['step_size'])
    out['seeds'] = np.array(case['seeds'])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, {k: v.shape for k, v in out.items() if k.startswith('samples')})


def run_reversibility(ref):
    """tests/test_util.py:97-110 -- the only hot-path test the reference owns."""
    from hamiltorch_b200 import targets as T
    var = torch.tensor([.10, .10])
    tgt = T.GaussianDiag(torch.zeros(2), var)

    def ref_log_prob(omega):     # the reference test's own closure
        return torch.distributions.MultivariateNormal(torch.zeros(2), torch.diag(var)).log_prob(omega).sum()

    q0 = torch.tensor([1., 1.])
    p0 = torch.tensor([1., 1.])
    inv_mass = torch.tensor([1., 1.])
    kw = dict(steps=100, step_size=0.1, inv_mass=inv_mass, sampler=ref.Sampler.HMC,
              integrator=ref.Integrator.EXPLICIT)
    out = {}
    for tag, lp in (('desc', tgt), ('mvn', ref_log_prob)):
        qf, pf = ref.samplers.leapfrog(q0, p0, lp, **kw)
        qb, pb = ref.samplers.leapfrog(qf[-1], -pf[-1].clone(), lp, **kw)
        out['fwd_q_' + tag] = torch.stack(qf).numpy()
        out['fwd_p_' + tag] = torch.stack(pf).numpy()
        out['bwd_q_' + tag] = torch.stack(qb).numpy()
        out['bwd_p_' + tag] = torch.stack(pb).numpy()
        print('reversibility[%s]: returns to start bitwise = %s' % (tag, bool(torch.all(qb[-1] == q0))))
    # oracle restatement agrees with the reference on the descriptor target
    oq, op = O.leapfrog_hmc(tgt, q0, p0, 100, 0.1, inv_mass)
    assert np.array_equal(torch.stack(oq).numpy(), out['fwd_q_desc'])
    assert np.array_equal(torch.stack(op).numpy(), out['fwd_p_desc'])
    np.savez_compressed(os.path.join(OUT, 'ref_reversibility.npz'), **out)


SCHEME_ID = {None: None, 'SPLITTING': O.SPLIT_SYM, 'SPLITTING_RAND': O.SPLIT_RAND, 'SPLITTING_KMID': O.SPLIT_KMID}


def build_mlp_case(case):
    """(model, x, y, descriptors, inv_mass, tau_list) of a case -- shared with the tests."""
    from hamiltorch_b200 import targets as T
    model, x, y = cases.mlp_problem(**case['problem'])
    tau_list = case.get('tau_list')
    tau_t = None if tau_list is None else [torch.tensor(t) for t in tau_list]
    loss = case.get('model_loss', 'regression')
    if case['scheme'] is None:
        descs = T.MLPTarget.from_model(model, x, y, tau_t, case['tau_out'], model_loss=loss)
    else:
        M = case['num_splits']
        B = x.shape[0] // M
        descs = [T.MLPTarget.from_model(model, x[m * B:(m + 1) * B], y[m * B:(m + 1) * B], tau_t, case['tau_out'],
                                        prior_scale=M, model_loss=loss) for m in range(M)]
    D = sum(p.numel() for p in model.parameters())
    inv_mass = None
    if case.get('diag_mass'):
        inv_mass = 0.5 + torch.rand(D, generator=torch.Generator().manual_seed(99))
    return model, x, y, descs, inv_mass, tau_t


def run_mlp_case(ref, name, case):
    import torch.utils.data as tud
    model, x, y, descs, inv_mass, tau_t = build_mlp_case(case)
    S, L, burn, eps = case['num_samples'], case['num_steps_per_sample'], case['burn'], case['step_size']
    D = sum(p.numel() for p in model.parameters())
    scheme = case['scheme']
    M = case.get('num_splits', 0)
    loss = case.get('model_loss', 'regression')
    out = {}
    for ci, seed in enumerate(case['seeds']):
        torch.manual_seed(seed)
        init = ref.util.flatten(model).detach().clone() + 0.05 * torch.randn(D)
        if scheme is None:
            samples = ref.sample_model(model, x, y, params_init=init, model_loss=loss, num_samples=S,
                                       num_steps_per_sample=L, step_size=eps, burn=burn, inv_mass=inv_mass,
                                       tau_out=case['tau_out'], tau_list=tau_t, verbose=False)
        else:
            loader = tud.DataLoader(tud.TensorDataset(x, y), batch_size=x.shape[0] // M, shuffle=False)
            samples = ref.sample_split_model(model, loader, params_init=init, num_splits=M, model_loss=loss,
                                             num_samples=S, num_steps_per_sample=L, step_size=eps, burn=burn,
                                             inv_mass=inv_mass, tau_out=case['tau_out'], tau_list=tau_t,
                                             integrator=getattr(ref.Integrator, scheme), verbose=False)
        samples = torch.stack(samples)
        # oracle replay from the same RNG state (descriptors instead of the reference's closures).  NB: iterating a
        # DataLoader draws its base seed from the global CPU generator (torch/utils/data/dataloader.py), which
        # define_split_model_log_prob does once (samplers.py:1251) -- part of the stream the reference consumes.
        torch.manual_seed(seed)
        init2 = ref.util.flatten(model).detach().clone() + 0.05 * torch.randn(D)
        if scheme is not None:
            next(iter(loader))
        res = O.sample_hmc(descs, init2, num_samples=S, num_steps_per_sample=L, step_size=eps, burn=burn,
                           inv_mass=inv_mass, split_scheme=SCHEME_ID[scheme])
        assert torch.equal(torch.stack(res['samples']), samples), name
        # the stream
        torch.manual_seed(seed)
        _ = 0.05 * torch.randn(D)
        if scheme is not None:
            next(iter(loader))
        nperm = M if scheme == 'SPLITTING_RAND' else 0
        z = torch.empty(S, D)
        logu = torch.empty(S)
        perms = torch.zeros(S, max(nperm, 1), dtype=torch.long)
        for n in range(S):
            z[n] = torch.randn(D)
            if nperm:
                perms[n] = torch.randperm(nperm)
            logu[n] = torch.log(torch.rand(1))[0]
        res2 = O.sample_hmc(descs, init2, num_samples=S, num_steps_per_sample=L, step_size=eps, burn=burn,
                            inv_mass=inv_mass, split_scheme=SCHEME_ID[scheme], normals=z, log_uniforms=logu,
                            perms=perms if nperm else None)
        assert torch.equal(torch.stack(res2['samples']), samples), name
        assert not any(res['diverged'])
        out['init_%d' % ci] = init.numpy()
        out['samples_%d' % ci] = samples.numpy()
        out['z_%d' % ci] = z.numpy()
        out['logu_%d' % ci] = logu.numpy()
        out['perms_%d' % ci] = perms.numpy()
        out['accepted_%d' % ci] = np.array(res['accepted'], dtype=np.uint8)
        out['ham_old_%d' % ci] = np.array(res['ham_old'], dtype=np.float64)
        out['ham_new_%d' % ci] = np.array(res['ham_new'], dtype=np.float64)
        if ci == 0:
            # predict_model on the retained samples (samplers.py:1468-1562)
            pred, lps = ref.predict_model(model, list(samples), x=x, y=y, model_loss=loss,
                                          tau_out=case['tau_out'], tau_list=tau_t)
            out['pred'] = pred.numpy()
            out['pred_log_prob'] = torch.stack([l.reshape(-1) for l in lps]).numpy()
    out['seeds'] = np.array(case['seeds'])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, out['samples_0'].shape, 'acc', [out['accepted_%d' % c].mean() for c in range(len(case['seeds']))])


def run_rmhmc_case(ref, name, case):
    tgt, D = case['target'], case['target'].dim
    S, L = case['num_samples'], case['num_steps_per_sample']
    explicit = case['integrator'] == 'EXPLICIT'
    kw = dict(num_samples=S, num_steps_per_sample=L, step_size=case['step_size'], burn=case['burn'],
              jitter=case['jitter'], softabs_const=case['softabs_const'])
    if explicit:
        kw['explicit_binding_const'] = case['explicit_binding_const']
    else:
        kw.update(fixed_point_threshold=case['fixed_point_threshold'],
                  fixed_point_max_iterations=case['fixed_point_max_iterations'])
    out = {}
    for ci, seed in enumerate(case['seeds']):
        init = torch.tensor(case['init'])
        torch.manual_seed(seed)
        samples = ref.sample(log_prob_func=tgt, params_init=init, sampler=ref.Sampler.RMHMC,
                             integrator=getattr(ref.Integrator, case['integrator']),
                             metric=getattr(ref.Metric, case['metric']), verbose=False, **kw)
        samples = torch.stack(samples)
        okw = dict(kw, integrator=R.EXPLICIT if explicit else R.IMPLICIT,
                   metric={'SOFTABS': R.SOFTABS, 'HESSIAN': R.HESSIAN, 'JACOBIAN_DIAG': R.JACOBIAN_DIAG}[case['metric']])
        torch.manual_seed(seed)
        res = R.sample_rmhmc(tgt, init, **okw)
        assert torch.equal(torch.stack(res['samples']), samples), name
        # the consumed stream: per iteration jitter(gibbs), normals, jitter(ham), 8L jitters (explicit), jitter(new_ham),
        # rand(1).  A LogProbError cuts the iteration short (fewer jitter draws, no rand(1)), NaN-gradient retries add
        # draws: the oracle's per-iteration counters say how many of each the reference consumed.  Recorded when the
        # count is knowable: explicit integrator, or no jitter at all.
        jittered = case['jitter'] is not None
        assert explicit or not jittered
        J = max(max(res['jitter_draws']), 8 * L + 3) if jittered else 0
        torch.manual_seed(seed)
        z = torch.zeros(S, D)
        logu = torch.zeros(S)
        uni = torch.zeros(S, max(J, 1), D)
        for n in range(S):
            k = res['jitter_draws'][n] if jittered else 0
            if k >= 1:
                uni[n, 0] = torch.rand(D)
            if res['gibbs_done'][n]:
                z[n] = torch.randn(D)
            for j in range(1, k):
                uni[n, j] = torch.rand(D)
            if not res['diverged'][n]:
                logu[n] = torch.log(torch.rand(1))[0]
        res2 = R.sample_rmhmc(tgt, init, normals=z, log_uniforms=logu, uniforms=uni if J else None, **okw)
        assert torch.equal(torch.stack(res2['samples']), samples), name + ' (injected)'
        out['samples_%d' % ci] = samples.numpy()
        out['z_%d' % ci] = z.numpy()
        out['logu_%d' % ci] = logu.numpy()
        out['uniforms_%d' % ci] = uni.numpy()
        out['accepted_%d' % ci] = np.array(res['accepted'], dtype=np.uint8)
        out['ham_old_%d' % ci] = np.array(res['ham_old'], dtype=np.float64)
        out['ham_new_%d' % ci] = np.array(res['ham_new'], dtype=np.float64)
        out['diverged_%d' % ci] = np.array(res['diverged'], dtype=np.uint8)
    out['seeds'] = np.array(case['seeds'])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('wrote', name, out['samples_0'].shape, 'acc', [float(out['accepted_%d' % c].mean()) for c in range(len(case['seeds']))],
          'LogProbError', [int(out['diverged_%d' % c].sum()) for c in range(len(case['seeds']))])


def run_split_standalone(ref):
    """Reference outputs of samplers.leapfrog called directly with a SPLITTING integrator (:494-603) on the descriptor list
    of the split Bayesian-NN cases: ret_params / ret_momenta after every step; the randperm(M) of SPLITTING_RAND (:550) is
    recorded."""
    out = {}
    for name in ('mlp_split_sym', 'mlp_split_rand', 'mlp_split_kmid', 'mlp_deep_tanh_mass'):
        case = cases.mlp_cases()[name]
        model, x, y, descs, inv_mass, tau_t = build_mlp_case(case)
        D = sum(p.numel() for p in model.parameters())
        M = case['num_splits']
        torch.manual_seed(41)
        q = ref.util.flatten(model).detach().clone() + 0.05 * torch.randn(D)
        p = torch.randn(D)
        st = torch.get_rng_state()
        qs, ps = ref.samplers.leapfrog(q.clone().requires_grad_(), p.clone(), descs, steps=3, step_size=case['step_size'],
                                       inv_mass=inv_mass, sampler=ref.Sampler.HMC,
                                       integrator=getattr(ref.Integrator, case['scheme']))
        torch.set_rng_state(st)
        perm = torch.randperm(M)
        oq, op = O.leapfrog_split(descs, q, p, 3, case['step_size'], inv_mass, SCHEME_ID[case['scheme']],
                                  perm if case['scheme'] == 'SPLITTING_RAND' else None)
        assert torch.equal(torch.stack([t.detach() for t in qs]), torch.stack(oq)), name
        out[name + '.q0'], out[name + '.p0'], out[name + '.perm'] = q.numpy(), p.numpy(), perm.numpy()
        out[name + '.q_traj'] = torch.stack([t.detach() for t in qs]).numpy()
        out[name + '.p_traj'] = torch.stack([t.detach() for t in ps]).numpy()
        print('split standalone', name, out[name + '.q_traj'].shape)
    np.savez_compressed(os.path.join(OUT, 'split_standalone.npz'), **out)


def run_rm_standalone(ref):
    """Reference outputs of leapfrog(sampler=RMHMC) -- (ret_params, ret_momenta[, params_copy, momentum_copy]) -- and of
    hamiltonian(sampler=RMHMC) at the start point; the jitter stream torch.rand(D) of the call is recorded."""
    out = {}
    for name, c in cases.standalone_rm_cases().items():
        tgt, D = c['target'], c['target'].dim
        q = torch.tensor(c['q'])
        torch.manual_seed(c['seed'])
        p = torch.randn(D)
        explicit = c['integrator'] == 'EXPLICIT'
        kw = dict(jitter=c['jitter'], softabs_const=c['softabs_const'], sampler=ref.Sampler.RMHMC,
                  integrator=getattr(ref.Integrator, c['integrator']), metric=getattr(ref.Metric, c['metric']))
        st = torch.get_rng_state()
        H = ref.samplers.hamiltonian(q.clone().requires_grad_(), p, tgt, explicit_binding_const=c.get('explicit_binding_const', 100),
                            **kw)
        torch.set_rng_state(st)
        out[name + '.uni_h'] = torch.rand(1, D).numpy()
        st = torch.get_rng_state()
        lk = dict(kw, steps=c['steps'], step_size=c['step_size'])
        if explicit:
            lk['explicit_binding_const'] = c['explicit_binding_const']
        ret_q, ret_p = ref.samplers.leapfrog(q.clone().requires_grad_(), p, tgt, **lk)
        torch.set_rng_state(st)
        J = 8 * c['steps'] if (explicit and c['jitter'] is not None) else 1
        out[name + '.uni_l'] = torch.rand(J, D).numpy()
        if explicit:
            (qs, qc), (ps, pc) = ret_q, ret_p
            out[name + '.q_copy'] = qc.detach().numpy()
            out[name + '.p_copy'] = pc.detach().numpy()
        else:
            qs, ps = ret_q, ret_p
        out[name + '.p0'] = p.numpy()
        out[name + '.H'] = H.detach().numpy().reshape(-1)
        out[name + '.q_traj'] = torch.stack([t.detach() for t in qs]).numpy()
        out[name + '.p_traj'] = torch.stack([t.detach() for t in ps]).numpy()
        print('standalone', name, 'H', float(H), 'q_L', out[name + '.q_traj'][-1][:3])
    np.savez_compressed(os.path.join(OUT, 'rmhmc_standalone.npz'), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    ref = import_reference()
    args = sys.argv[1:]
    only = {a[5:] for a in args if a.startswith('only:')}          # only:<case name> regenerates just those fixtures
    which = [a for a in args if not a.startswith('only:')] or ['plain', 'mlp', 'rmhmc']
    if not only:
        run_reversibility(ref)
    if 'plain' in which:
        for name, case in cases.plain_cases().items():
            if not only or name in only:
                run_plain_case(ref, name, case)
    if 'mlp' in which:
        for name, case in cases.mlp_cases().items():
            if not only or name in only:
                run_mlp_case(ref, name, case)
    if 'rmhmc' in which:
        for name, case in cases.rmhmc_cases().items():
            if not only or name in only:
                run_rmhmc_case(ref, name, case)
        if not only or 'rmhmc_standalone' in only:
            run_rm_standalone(ref)
    if 'mlp' in which and (not only or 'split_standalone' in only):
        run_split_standalone(ref)


if __name__ == '__main__':
    main()
