"""Developer tool: clock64() stamps (thread 0 of CTA 0) through ONE leapfrog step of the config-4 run kernel.
Build with  HMCX_NVCC_EXTRA=-DHMCX_TC_PROF python -m hamiltorch_b200.build --force  first."""
import ctypes as C, sys
sys.path.insert(0, '.')
import torch
import torch.nn as nn
import hamiltorch_b200 as hb
from hamiltorch_b200 import targets as T, _native as N

Cn = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator().manual_seed(0)
X = torch.randn(1024, 64, generator=g); w = torch.randn(64, 1, generator=g)
y = torch.sin(X @ w / 8) + 0.1 * torch.randn(1024, 1, generator=g)
torch.manual_seed(0)
model = nn.Sequential(nn.Linear(64, 128), nn.ReLU(), nn.Linear(128, 1))
descs = [T.MLPRegression.from_model(model, X[m * 256:(m + 1) * 256], y[m * 256:(m + 1) * 256], None, 100., prior_scale=4) for m in range(4)]
D = descs[0].dim
init = (hb.util.flatten(model).detach()[None] + 0.01 * torch.randn(Cn, D, generator=g)).cuda()
lib = N.load_library()
lib.hmcx_debug_tc_prof.restype = C.c_int
buf = (C.c_longlong * 512)()
for rep in range(2):
    hb.sample_chains(descs, init, num_samples=2, num_steps_per_sample=1, step_size=5e-4, inv_mass=torch.ones(D),
                     integrator=hb.Integrator.SPLITTING, rng='philox', seed=3)
    torch.cuda.synchronize()
    n = lib.hmcx_debug_tc_prof(buf)
    ev = [(buf[i] >> 48, buf[i] & 0xFFFFFFFFFFFF) for i in range(n)]
    print('rep', rep, 'marks', n)
    print(' '.join('%d:+%d' % (ev[i][0], ev[i][1] - ev[i - 1][1]) for i in range(1, n)))
