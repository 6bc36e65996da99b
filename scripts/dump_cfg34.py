"""Debug helper (GPU box): dump the kernel's teacher-forced outputs for configs 3 and 4 next to the oracle's into
gpurun_out/ for analysis in the build container."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hamiltorch_b200 import engine, targets as T, util, _native as N       # noqa: E402
from oracle import gen_cfg3 as G, cfg4                                      # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
os.makedirs(OUT, exist_ok=True)


def cfg3():
    d = np.load('tests/golden/cfg3_rmhmc_tf.npz')
    C, S = d['accepted'].shape
    zs, lus, us = zip(*[G.stream(c, S) for c in range(C)])
    z, logu, uni = torch.stack(zs, 1), torch.stack(lus, 1), torch.stack(us, 1)
    init = torch.from_numpy(d['state_in']).reshape(C * S, 2)
    kw = dict(burn=0, jitter=G.JIT, softabs_const=G.ALPHA, explicit_binding_const=G.OMEGA, explicit=True, softabs=True,
              record_ham=True)
    res = engine.rmhmc_run(T.Funnel(2), init, 1, G.L, G.EPS, normals=z.permute(1, 0, 2).reshape(1, C * S, 2),
                           log_uniforms=logu.permute(1, 0).reshape(1, C * S),
                           uniforms=uni.permute(1, 0, 2, 3).reshape(1, C * S, G.J, 2), **kw)
    torch.cuda.synchronize()
    free = engine.rmhmc_run(T.Funnel(2), torch.tensor(G.INIT).repeat(C, 1), S, G.L, G.EPS, normals=z, log_uniforms=logu,
                            uniforms=uni, **kw)
    torch.cuda.synchronize()
    np.savez_compressed(os.path.join(OUT, 'dump_cfg3.npz'),
                        acc=res.accepted[:, 0].cpu().numpy().reshape(C, S), div=res.diverged[:, 0].cpu().numpy().reshape(C, S),
                        ham=res.ham[:, 0].cpu().numpy().reshape(C, S, 2), q=res.final_state.cpu().numpy().reshape(C, S, 2),
                        free_acc=free.accepted.cpu().numpy(), free_div=free.diverged.cpu().numpy(),
                        free_samples=free.samples.cpu().numpy(), free_ham=free.ham.cpu().numpy())
    # timing of config 3 at BASELINE size
    import hamiltorch_b200 as hb
    init3 = torch.tensor([0., 1.], device='cuda').repeat(512, 1)
    f = lambda: hb.sample_chains(T.Funnel(2), init3, num_samples=200, num_steps_per_sample=10, step_size=0.05, jitter=1e-3,
                                 softabs_const=1e6, explicit_binding_const=10, sampler=hb.Sampler.RMHMC,
                                 integrator=hb.Integrator.EXPLICIT, metric=hb.Metric.SOFTABS, rng='philox', seed=2)
    r = f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = f(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print('config3: %.3f ms, %.3e chain-steps/s, acc %.3f, LogProbError %.3f' % (
        ms, 512 * 200 * 10 / ms * 1e3, float(r.accepted.float().mean()), float(r.diverged.float().mean())))


def cfg4_dump():
    import multiprocessing as mp
    C, S = 8, 100
    t0 = time.time()
    with mp.get_context('spawn').Pool(2 * C) as pool:
        rows = pool.map(cfg4.run_chain, [(c, S) for c in range(C)])
    print('oracle %.1f s' % (time.time() - t0))
    smp_ref = np.stack([r[0] for r in rows]); acc_ref = np.stack([r[1] for r in rows]); ham_ref = np.stack([r[2] for r in rows])
    model, X, y = cfg4.problem()
    descs = cfg4.descriptors(model, X, y)
    D = descs[0].dim
    flat = util.flatten(model).detach().clone()
    st = [cfg4.stream(c, S, D, flat) for c in range(C)]
    init, z, logu = torch.stack([s[0] for s in st]), torch.stack([s[1] for s in st], 1), torch.stack([s[2] for s in st], 1)
    out = {}
    for tc in (0, 1):
        for d_ in descs:
            d_.tensor_cores = tc
        res = engine.hmc_run(descs, init, S, cfg4.L, cfg4.EPS, inv_mass=torch.ones(D), normals=z, log_uniforms=logu,
                             record_ham=True, scheme=N.SCHEME_SPLIT_SYM)
        torch.cuda.synchronize()
        its = np.arange(2, S)
        q_in = torch.from_numpy(smp_ref[:, its - 1]).reshape(-1, D)
        tf = engine.hmc_run(descs, q_in, 1, cfg4.L, cfg4.EPS, inv_mass=torch.ones(D),
                            normals=z[its].permute(1, 0, 2).reshape(1, -1, D),
                            log_uniforms=logu[its].permute(1, 0).reshape(1, -1), record_ham=True,
                            scheme=N.SCHEME_SPLIT_SYM)
        torch.cuda.synchronize()
        q = tf.final_state.cpu().numpy().reshape(C, -1, D)
        err = np.abs(q - smp_ref[:, its]).max(-1)
        free_err = np.abs(res.samples.cpu().numpy() - smp_ref).max(-1)
        out['tf_err_%d' % tc] = err
        out['tf_acc_%d' % tc] = tf.accepted[:, 0].cpu().numpy().reshape(C, -1)
        out['free_err_%d' % tc] = free_err
        out['free_acc_%d' % tc] = res.accepted.cpu().numpy()
        out['free_ham_%d' % tc] = res.ham.cpu().numpy()
        out['tf_ham_%d' % tc] = tf.ham[:, 0].cpu().numpy().reshape(C, -1, 2)
        print('tc_off=%d: TF err max %.2e median %.2e; free err max %.2e' % (tc, err.max(), np.median(err), free_err.max()))
    np.savez_compressed(os.path.join(OUT, 'dump_cfg4.npz'), acc_ref=acc_ref, ham_ref=ham_ref, logu=logu.numpy(),
                        scale=np.abs(smp_ref).max(), **out)


if __name__ == '__main__':
    which = sys.argv[1:] or ['3', '4']
    if '3' in which:
        cfg3()
    if '4' in which:
        cfg4_dump()
