"""A/B of programmatic dependent launch on the step-synchronous dense path (run once per HMCX_PDL setting)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hamiltorch_b200 as hb
from hamiltorch_b200 import targets as T, engine

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for C, D in ((256, 1024), (1024, 2048), (512, 64)):
    g = torch.Generator().manual_seed(3)
    A = torch.randn(D, D, generator=g, dtype=torch.float64) / D ** 0.5
    tgt = engine.NativeTarget(T.GaussianFull(torch.zeros(D), cov=A @ A.t() + 0.5 * torch.eye(D, dtype=torch.float64)), 'cuda')
    init = torch.randn(C, D, generator=g).cuda()
    S, L = 40, 10
    ms = timed(lambda: engine.hmc_run(tgt, init, S, L, 0.1, seed=5))
    print(json.dumps(dict(pdl=os.environ.get('HMCX_PDL', '1'), C=C, D=D, ms=ms, us_per_step_launch=1e3 * ms / (S * (L + 1)),
                          chain_steps_per_s=C * S * L / (ms * 1e-3))), flush=True)
