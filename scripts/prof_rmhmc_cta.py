"""ncu target: the CTA-per-chain RMHMC kernel (in-kernel metric assembly + Jacobi eigensolver) on the D=32 funnel.
    ncu --set full --clock-control none --import-source on -k regex:rmhmc_cta_kernel -c 1 -o gpurun_out/r2_rmcta \
        python scripts/prof_rmhmc_cta.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                     # noqa: E402
import hamiltorch_b200 as hb                     # noqa: E402
from hamiltorch_b200 import targets as T         # noqa: E402

C, D, S, L = int(os.environ.get('C', '592')), int(os.environ.get('D', '32')), int(os.environ.get('S', '2')), int(os.environ.get('L', '3'))
tgt = T.Funnel(D)
init = (torch.ones(C, D) * 0.5).cuda()
init[:, 0] = 0.0
res = hb.sample_chains(tgt, init, num_samples=S, num_steps_per_sample=L, step_size=0.03, sampler=hb.Sampler.RMHMC,
                       integrator=hb.Integrator.EXPLICIT, metric=hb.Metric.SOFTABS, jitter=1e-3, softabs_const=1e6,
                       explicit_binding_const=10, rng='philox', seed=5)
torch.cuda.synchronize()
print('ok accept', float(res.accepted.float().mean()))
