"""ncu target: dense_step_kernel (GaussianFull plain HMC), env C / D pick the shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamiltorch_b200 as hb
from hamiltorch_b200 import targets as T, engine
C, D = int(os.environ.get('C', '1024')), int(os.environ.get('D', '2048'))
g = torch.Generator().manual_seed(3)
A = torch.randn(D, D, generator=g, dtype=torch.float64) / D ** 0.5
tgt = engine.NativeTarget(T.GaussianFull(torch.zeros(D), cov=A @ A.t() + 0.5 * torch.eye(D, dtype=torch.float64)), 'cuda')
init = torch.randn(C, D, generator=g).cuda()
for _ in range(2):
    r = engine.hmc_run(tgt, init, 3, 4, 0.1, seed=5)
    torch.cuda.synchronize()
print('ok accept', float(r.accepted.float().mean()))
