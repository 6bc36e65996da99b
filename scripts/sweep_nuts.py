"""Persistent kernel: NUTS on/off x dimension (CTA size) x geometry.  us per iteration."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hamiltorch_b200 import engine, targets as T

C, S, L, burn = 128, 150, 10, 100
only = os.environ.get('ONLY')
for D in (1024, 2048, 3072, 4096):
    tgt = engine.NativeTarget(T.GaussianIso(D), 'cuda')
    init = (0.1 * torch.randn(C, D, generator=torch.Generator().manual_seed(0))).cuda()
    for nuts in (True, False):
        for tuning in ((0, 2) if D > 1024 else (0,)):
            if only and only != '%d-%d-%d' % (D, int(nuts), tuning):
                continue
            def fn():
                return engine.hmc_run(tgt, init, S, L, 0.1, burn=burn if nuts else 0, nuts=nuts, seed=1, tuning=tuning)
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                r = fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(json.dumps(dict(D=D, nuts=nuts, tuning=tuning, ms=round(ms, 3), us_per_iteration=round(1e3 * ms / S, 2),
                                  accept=round(float(r.accepted.float().mean()), 3))), flush=True)
