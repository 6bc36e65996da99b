"""Diagnostic: per-launch CUDA-event times of the config-2 launch, 40 launches back to back, with SM clock samples."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hamiltorch_b200 import engine, targets as T, _native as N
dev = torch.device('cuda', 0)
tgt = engine.NativeTarget(T.GaussianIso(1024), dev)
q0 = (0.1 * torch.randn(256, 1024)).to(dev)
out = torch.empty((256, 1000, 1024), dtype=torch.float32, device=dev)
def smi():
    return subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active',
                           '--format=csv,noheader'], capture_output=True, text=True).stdout.strip()
print('before', smi())
evs = []
for k in range(40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = engine.hmc_run(tgt, q0, 1000, 10, 0.05, seed=k, out=out, device=dev)
    e1.record()
    evs.append((e0, e1))
    if k in (5, 20):
        torch.cuda.synchronize(); print('during', smi())
torch.cuda.synchronize()
print(' '.join('%.2f' % a.elapsed_time(b) for a, b in evs))
print('after', smi())
