"""ncu target: the persistent small-D flow kernel (hmcx_flow.cu) on constant-metric explicit RMHMC, env C / D / S pick the shape.
    ncu --set full --clock-control none --import-source on -k regex:flow_small_kernel -c 1 -o gpurun_out/r2_flow \
        python scripts/prof_flow.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                     # noqa: E402
import hamiltorch_b200 as hb                     # noqa: E402
from hamiltorch_b200 import targets as T         # noqa: E402

C, D, S = int(os.environ.get('C', '512')), int(os.environ.get('D', '64')), int(os.environ.get('S', '50'))
g = torch.Generator().manual_seed(3)
A = torch.randn(D, D, generator=g, dtype=torch.float64) / D ** 0.5
tgt = T.GaussianFull(torch.zeros(D), cov=A @ A.t() + 0.5 * torch.eye(D, dtype=torch.float64))
init = (torch.randn(C, D, generator=g) * 0.5).cuda()
res = hb.sample_chains(tgt, init, num_samples=S, num_steps_per_sample=10, step_size=0.1, explicit_binding_const=10,
                       sampler=hb.Sampler.RMHMC, integrator=hb.Integrator.EXPLICIT, metric=hb.Metric.HESSIAN,
                       rng='philox', seed=2)
torch.cuda.synchronize()
print('ok accept', float(res.accepted.float().mean()))
