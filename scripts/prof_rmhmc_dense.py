"""ncu target: the constant-metric RMHMC path (hmcx_rmhmc_dense_run), env C / D pick the shape.
    ncu --set full --clock-control none --import-source on -k regex:dense_lin_kernel -s 30 -c 3 -o gpurun_out/prof_lin \
        python scripts/prof_rmhmc_dense.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                     # noqa: E402
import hamiltorch_b200 as hb                     # noqa: E402
from hamiltorch_b200 import targets as T         # noqa: E402

C, D = int(os.environ.get('C', '1024')), int(os.environ.get('D', '1024'))
g = torch.Generator().manual_seed(3)
A = torch.randn(D, D, generator=g, dtype=torch.float64) / D ** 0.5
tgt = T.GaussianFull(torch.zeros(D), cov=A @ A.t() + 0.5 * torch.eye(D, dtype=torch.float64))
init = torch.randn(C, D, generator=g).cuda()
for _ in range(2):
    res = hb.sample_chains(tgt, init, num_samples=3, num_steps_per_sample=4, step_size=0.1, explicit_binding_const=10,
                           sampler=hb.Sampler.RMHMC, integrator=hb.Integrator.EXPLICIT, metric=hb.Metric.HESSIAN,
                           rng='philox', seed=2)
    torch.cuda.synchronize()
print('ok accept', float(res.accepted.float().mean()))
