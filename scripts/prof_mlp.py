import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamiltorch_b200 as hb
from hamiltorch_b200 import targets as T
from oracle import cases
model, x, y = cases.mlp_problem(seed=0, n=1024, n_in=64, hidden=128)
M = 4
descs = [T.MLPRegression.from_model(model, x[m * 256:(m + 1) * 256], y[m * 256:(m + 1) * 256], None, 100., prior_scale=M) for m in range(M)]
D = descs[0].dim
C = int(os.environ.get('C', '64'))
init = hb.util.flatten(model).detach()[None] + 0.01 * torch.randn(C, D, generator=torch.Generator().manual_seed(0))
for _ in range(2):
    res = hb.sample_chains(descs, init.cuda(), num_samples=2, num_steps_per_sample=5, step_size=5e-4, inv_mass=torch.ones(D),
                           integrator=hb.Integrator.SPLITTING, rng='philox', seed=3)
    torch.cuda.synchronize()
print('ok')
