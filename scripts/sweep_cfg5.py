"""Geometry sweep of the persistent kernel on BASELINE config 5's per-GPU share (128 chains, D=4096, HMC_NUTS)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hamiltorch_b200 import engine, targets as T

C, D, S, L, burn = 128, 4096, 150, 10, 100
tgt = engine.NativeTarget(T.GaussianIso(D), 'cuda')
init = (0.1 * torch.randn(C, D, generator=torch.Generator().manual_seed(0))).cuda()
for nuts in (True, False):
    for tuning in (0, 2, 4):
        def fn():
            return engine.hmc_run(tgt, init, S, L, 0.1, burn=burn if nuts else 0, nuts=nuts, seed=1, tuning=tuning)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            r = fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps(dict(nuts=nuts, tuning=tuning, ms=ms, chain_steps_per_s=C * S * L / (ms * 1e-3),
                              us_per_iteration=1e3 * ms / S, accept=float(r.accepted.float().mean()))), flush=True)
