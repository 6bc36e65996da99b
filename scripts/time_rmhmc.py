"""Device time of the in-kernel-metric RMHMC paths (thread-per-chain D <= 16, CTA-per-chain 16 < D <= 64), Philox RNG.
    python scripts/time_rmhmc.py"""
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


def main():
    dev = torch.device('cuda', 0)
    cases = [
        ('funnel D=5 explicit softabs jitter (thread/chain)', T.Funnel(5), 4096, 20, 5, 0.05, dict(integrator=hb.Integrator.EXPLICIT, metric=hb.Metric.SOFTABS, jitter=1e-3, softabs_const=1e6, explicit_binding_const=10)),
        ('funnel D=11 explicit softabs jitter (thread/chain)', T.Funnel(11), 4096, 10, 5, 0.05, dict(integrator=hb.Integrator.EXPLICIT, metric=hb.Metric.SOFTABS, jitter=1e-3, softabs_const=1e6, explicit_binding_const=10)),
        ('funnel D=5 implicit softabs (thread/chain)', T.Funnel(5), 4096, 10, 5, 0.05, dict(integrator=hb.Integrator.IMPLICIT, metric=hb.Metric.SOFTABS, jitter=1e-3, softabs_const=1e6)),
        ('funnel D=32 explicit softabs jitter (CTA/chain)', T.Funnel(32), 592, 10, 5, 0.03, dict(integrator=hb.Integrator.EXPLICIT, metric=hb.Metric.SOFTABS, jitter=1e-3, softabs_const=1e6, explicit_binding_const=10)),
        ('diag D=64 explicit hessian jitter (CTA/chain)', T.GaussianDiag(torch.linspace(-1, 1, 64), 0.5 + torch.rand(64)), 592, 10, 5, 0.2, dict(integrator=hb.Integrator.EXPLICIT, metric=hb.Metric.HESSIAN, jitter=1e-3, explicit_binding_const=10)),
    ]
    for name, tgt, C, S, L, eps, kw in cases:
        D = tgt.dim
        init = (torch.ones(C, D) * 0.5).to(dev)
        if isinstance(tgt, T.Funnel):
            init[:, 0] = 0.0
        ms, r = timed(lambda: hb.sample_chains(tgt, init, num_samples=S, num_steps_per_sample=L, step_size=eps,
                                               sampler=hb.Sampler.RMHMC, rng='philox', seed=5, **kw))
        print('%-52s C=%d S=%d L=%d  %.2f ms  %.3g chain-steps/s  accept %.2f diverged %.2f'
              % (name, C, S, L, ms, C * S * L / (ms * 1e-3), float(r.accepted.float().mean()), float(r.diverged.float().mean())))


if __name__ == '__main__':
    main()
