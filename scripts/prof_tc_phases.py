"""Developer tool: per-phase clock64() stamps of ONE tensor-core gradient evaluation (thread 0 of CTA 0).
Build the library with  HMCX_NVCC_EXTRA=-DHMCX_TC_PROF python -m hamiltorch_b200.build --force  first."""
import ctypes as C, sys
sys.path.insert(0, '.')
import torch
import hamiltorch_b200 as hb
from hamiltorch_b200 import engine, targets as T, _native as N
from oracle import cases

model, x, y = cases.mlp_problem(seed=7, n=1024, n_in=64, hidden=128)
d = [T.MLPTarget.from_model(model, x[m*256:(m+1)*256], y[m*256:(m+1)*256], None, 100., prior_scale=4) for m in range(4)]
q = hb.util.flatten(model).detach()[None] + 0.05 * torch.randn(1, 8449)
lib = N.load_library()
lib.hmcx_debug_tc_prof.restype = C.c_int
buf = (C.c_longlong * 64)()
for rep in range(3):
    engine.grad_log_prob(d, q, split=1, want_log_prob=False)
    n = lib.hmcx_debug_tc_prof(buf)
    ev = [(buf[i] >> 48, buf[i] & 0xFFFFFFFFFFFF) for i in range(n)]
    print('rep', rep, ' '.join('%d:+%d' % (ev[i][0], ev[i][1] - ev[i-1][1]) for i in range(1, n)))
