"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Fixtures for BASELINE config 3 (explicit RMHMC, 2-D funnel, softabs 1e6,
omega=10, eps=.05, L=10, jitter=1e-3, init (0,1)) at the reference's full per-chain length.

    python -m oracle.gen_cfg3 [pin] [tf] [stats]          (build container; `pin` needs /root/reference)

Why three fixtures.  Config 3 is CHAOTIC in fp32: the reference's own chain, evaluated once in fp32 and once in
fp64 from the same random stream, makes its first different accept/reject decision after 11-68 iterations (median
~19) and is O(1) apart afterwards (tests/golden/cfg3_rmhmc_tf.npz records both).  About a quarter of its iterations
end in LogProbError and the NaN-retry loop of samplers.py:402-410 fires regularly.  No fp32 implementation -- the
reference on another CPU included -- reproduces a 100-iteration chain of it to 1e-4, so parity is decomposed:

  cfg3_rmhmc_pin.npz    the UNMODIFIED reference under torch's global RNG on chains that reject, raise LogProbError
                        and retry NaN gradients; the oracle is asserted bit-identical (pins those paths).
  cfg3_rmhmc_tf.npz     8 chains x 100 iterations from an injected stream: every iteration's input state, proposal,
                        Hamiltonians, decision and LogProbError flag from the fp32 oracle, and the SAME transition
                        (from the same input state) evaluated in fp64 -> the per-transition round-off floor of the
                        reference itself.  The GPU test restarts the kernel from every input state ("teacher
                        forcing", like the NUTS schedule) and must match decisions and stay within that floor.
  cfg3_rmhmc_stats.npz  64 chains x 200 iterations: pooled posterior moments, acceptance and LogProbError rates
                        with Monte-Carlo standard errors, for the statistical check of the free-running kernel at
                        BASELINE size (512 chains, in-kernel Philox).
"""
import multiprocessing as mp
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from hamiltorch_b200 import targets as T           # noqa: E402
from oracle import rmhmc_oracle as R               # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
L, EPS, JIT, ALPHA, OMEGA = 10, 0.05, 1e-3, 1e6, 10
J = 8 * L + 3 + 24                  # rows of jitter uniforms per iteration: the fixed 8L+3 plus room for NaN retries
KW = dict(num_steps_per_sample=L, step_size=EPS, jitter=JIT, softabs_const=ALPHA, explicit_binding_const=OMEGA,
          integrator=R.EXPLICIT, metric=R.SOFTABS)
INIT = [0., 1.]


def stream(seed, S):
    """The injected random stream of one chain (CPU generator: identical on every machine with this torch)."""
    g = torch.Generator().manual_seed(7000 + seed)
    z = torch.randn(S, 2, generator=g)
    logu = torch.log(torch.rand(S, generator=g))
    uni = torch.rand(S, J, 2, generator=g)
    return z, logu, uni


def _quiet():
    sys.stdout = open(os.devnull, 'w')              # the reference prints 'Invalid ...' on every LogProbError
    torch.set_num_threads(1)
    import warnings
    warnings.filterwarnings('ignore')


# ---------------------------------------------------------------------------------------------------------------
def _pin_one(seed):
    _quiet()
    from oracle.ref_import import import_reference
    ref = import_reference()
    S = 25
    tgt, init = T.Funnel(2), torch.tensor(INIT)
    torch.manual_seed(seed)
    samples = ref.sample(log_prob_func=tgt, params_init=init, num_samples=S, num_steps_per_sample=L, step_size=EPS,
                         burn=3, jitter=JIT, softabs_const=ALPHA, explicit_binding_const=OMEGA,
                         sampler=ref.Sampler.RMHMC, integrator=ref.Integrator.EXPLICIT, metric=ref.Metric.SOFTABS,
                         verbose=False)
    torch.manual_seed(seed)
    res = R.sample_rmhmc(tgt, init, num_samples=S, burn=3, **KW)
    assert torch.equal(torch.stack(res['samples']), torch.stack(samples)), 'oracle != reference at seed %d' % seed
    return (seed, torch.stack(samples).numpy(), np.array(res['accepted'], np.uint8), np.array(res['diverged'], np.uint8),
            np.array(res['nan_retries'], np.int32))


def gen_pin():
    with mp.get_context('fork').Pool(8) as pool:
        rows = pool.map(_pin_one, range(40, 48))
    # keep the chains that together exercise reject, LogProbError and the NaN-retry loop
    rows.sort(key=lambda r: -(int(r[3].sum() > 0) + int((r[4] > 0).any()) + int(((r[2] == 0) & (r[3] == 0)).any())))
    rows = rows[:2]
    out = {'seeds': np.array([r[0] for r in rows])}
    for ci, (seed, smp, acc, div, jd) in enumerate(rows):
        out['samples_%d' % ci], out['accepted_%d' % ci], out['diverged_%d' % ci], out['nan_retries_%d' % ci] = \
            smp, acc, div, jd
        print('pin seed', seed, 'acc', acc.mean(), 'LogProbError', int(div.sum()), 'plain rejects',
              int(((acc == 0) & (div == 0)).sum()), 'NaN retries', int(jd.sum()), file=sys.stderr)
    assert any(r[3].sum() > 0 for r in rows) and any((r[4] > 0).any() for r in rows)
    assert any(((r[2] == 0) & (r[3] == 0)).any() for r in rows)
    np.savez_compressed(os.path.join(OUT, 'cfg3_rmhmc_pin.npz'), **out)


# ---------------------------------------------------------------------------------------------------------------
def _tf_one(seed):
    _quiet()
    S = 100
    tgt = T.Funnel(2)
    z, logu, uni = stream(seed, S)
    r = R.sample_rmhmc(tgt, torch.tensor(INIT), num_samples=S, burn=0, normals=z, log_uniforms=logu, uniforms=uni,
                       **KW)
    state = torch.stack(r['state_in'])
    prop64 = np.full((S, 2), np.nan)
    ham64 = np.full((S, 2), np.nan)
    acc64 = np.zeros(S, np.uint8)
    div64 = np.zeros(S, np.uint8)
    for n in range(S):                                   # the same transition in fp64, from the fp32 chain's state
        r1 = R.sample_rmhmc(tgt, state[n].double(), num_samples=1, burn=0, normals=z[n:n + 1].double(),
                            log_uniforms=logu[n:n + 1].double(), uniforms=uni[n:n + 1].double(), **KW)
        prop64[n] = r1['proposal'][0].numpy()
        ham64[n] = [r1['ham_old'][0], r1['ham_new'][0]]
        acc64[n], div64[n] = r1['accepted'][0], r1['diverged'][0]
    # and the free-running fp64 chain (how long do fp32 and fp64 stay together?)
    r64 = R.sample_rmhmc(tgt, torch.tensor(INIT, dtype=torch.float64), num_samples=S, burn=0, normals=z.double(),
                         log_uniforms=logu.double(), uniforms=uni.double(), **KW)
    return dict(state_in=state.numpy(), proposal=torch.stack(r['proposal']).numpy(),
                ham=np.stack([r['ham_old'], r['ham_new']], 1), accepted=np.array(r['accepted'], np.uint8),
                diverged=np.array(r['diverged'], np.uint8), jitter_draws=np.array(r['jitter_draws'], np.int32),
                nan_retries=np.array(r['nan_retries'], np.int32),
                samples=torch.stack(r['samples']).numpy(), proposal64=prop64, ham64=ham64, accepted64=acc64,
                diverged64=div64, free64_accepted=np.array(r64['accepted'], np.uint8),
                free64_samples=torch.stack(r64['samples']).numpy())


def gen_tf():
    seeds = list(range(8))
    with mp.get_context('fork').Pool(8) as pool:
        rows = pool.map(_tf_one, seeds)
    out = {k: np.stack([r[k] for r in rows]) for k in rows[0]}          # leading axis = chain
    out['seeds'] = np.array(seeds)
    np.savez_compressed(os.path.join(OUT, 'cfg3_rmhmc_tf.npz'), **out)
    a, d = out['accepted'], out['diverged']
    first = [int(np.argmax(x != y)) if (x != y).any() else len(x) for x, y in zip(a, out['free64_accepted'])]
    err = np.abs(out['proposal'] - out['proposal64'])
    ok = np.isfinite(err).all(-1)
    print('tf: acc %.3f LogProbError %.3f plain rejects %.3f; fp32 vs fp64 free chains first differ at %s; '
          'per-transition |prop32-prop64| median %.2e p90 %.2e max %.2e; decisions differ on %d of %d transitions'
          % (a.mean(), d.mean(), ((a == 0) & (d == 0)).mean(), first, np.median(err[ok]), np.percentile(err[ok], 90),
             err[ok].max(), int((a != out['accepted64']).sum()), a.size), file=sys.stderr)


# ---------------------------------------------------------------------------------------------------------------
def _stats_one(seed):
    _quiet()
    S = 200
    z, logu, uni = stream(1000 + seed, S)
    r = R.sample_rmhmc(T.Funnel(2), torch.tensor(INIT), num_samples=S, burn=0, normals=z, log_uniforms=logu,
                       uniforms=uni, **KW)
    return torch.stack(r['samples']).numpy(), np.array(r['accepted'], np.uint8), np.array(r['diverged'], np.uint8)


def gen_stats():
    C = 64
    with mp.get_context('fork').Pool(8) as pool:
        rows = pool.map(_stats_one, range(C))
    smp = np.stack([r[0] for r in rows])                    # (C, S, 2)
    acc = np.stack([r[1] for r in rows])
    div = np.stack([r[2] for r in rows])
    np.savez_compressed(os.path.join(OUT, 'cfg3_rmhmc_stats.npz'), samples=smp.astype(np.float32), accepted=acc,
                        diverged=div)
    half = smp[:, smp.shape[1] // 2:]
    print('stats: acc %.3f +- %.3f, LogProbError rate %.3f, v mean %.3f sd %.3f, x sd %.3f'
          % (acc.mean(), acc.mean(1).std() / C ** 0.5, div.mean(), half[..., 0].mean(), half[..., 0].std(),
             half[..., 1].std()), file=sys.stderr)


if __name__ == '__main__':
    which = sys.argv[1:] or ['pin', 'tf', 'stats']
    os.makedirs(OUT, exist_ok=True)
    if 'pin' in which:
        gen_pin()
    if 'tf' in which:
        gen_tf()
    if 'stats' in which:
        gen_stats()
