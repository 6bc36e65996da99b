"""Device time of the tcgen05 dense paths (no CPU legs): GaussianFull / full inv_mass / constant-metric RMHMC.
    python scripts/time_dense.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hamiltorch_b200 as hb                                  # noqa: E402
from hamiltorch_b200 import targets as T                      # noqa: E402


def timed(f, reps=3):
    f()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r


def spd(D, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(D, D, generator=g, dtype=torch.float64) / D ** 0.5
    return A @ A.t() + 0.5 * torch.eye(D, dtype=torch.float64)


def main():
    rows = []
    for C, D, S, L in ((256, 1024, 20, 10), (1024, 2048, 10, 10)):
        tgt = T.GaussianFull(torch.zeros(D), cov=spd(D, 3))
        init = torch.randn(C, D, generator=torch.Generator().manual_seed(5)).cuda()
        ms, r = timed(lambda: hb.sample_chains(tgt, init, num_samples=S, num_steps_per_sample=L, step_size=0.1, rng='philox', seed=5))
        fl = 2.0 * C * D * D * (L + 1) * S
        rows.append(('GaussianFull D=%d x %d chains (dense_step_kernel), L=%d S=%d' % (D, C, L, S), ms, C * S * L, fl, r))
    C, D, S, L = 256, 1024, 10, 10
    im = spd(D, 4).float()
    init = torch.randn(C, D, generator=torch.Generator().manual_seed(5)).cuda()
    ms, r = timed(lambda: hb.sample_chains(T.GaussianIso(D), init, num_samples=S, num_steps_per_sample=L, step_size=0.1, inv_mass=im, rng='philox', seed=5))
    rows.append(('full inv_mass, GaussianIso D=%d x %d chains (dense_lin_kernel), L=%d S=%d' % (D, C, L, S), ms, C * S * L, 2.0 * C * D * D * (L + 3) * S, r))
    for C, D, S, L in ((512, 64, 20, 10), (1024, 1024, 4, 10)):
        tgt = T.GaussianFull(torch.zeros(D), cov=spd(D, 3))
        init = torch.randn(C, D, generator=torch.Generator().manual_seed(5)).cuda() * 0.5
        ms, r = timed(lambda: hb.sample_chains(tgt, init, num_samples=S, num_steps_per_sample=L, step_size=0.1, sampler=hb.Sampler.RMHMC,
                                               integrator=hb.Integrator.EXPLICIT, metric=hb.Metric.HESSIAN, explicit_binding_const=10, rng='philox', seed=5))
        rows.append(('explicit RMHMC constant metric, GaussianFull D=%d x %d chains (dense_lin_kernel), L=%d S=%d' % (D, C, L, S), ms, C * S * L, 2.0 * C * D * D * (8 * L + 3) * S, r))
    for name, ms, steps, fl, r in rows:
        print('%-95s %8.2f ms  %.3g chain-steps/s  %.1f algorithmic TFLOP/s  accept %.2f' % (name, ms, steps / (ms * 1e-3), fl / (ms * 1e-3) / 1e12, float(r.accepted.float().mean())))


if __name__ == '__main__':
    main()
