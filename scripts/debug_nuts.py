import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hamiltorch_b200 import engine
from oracle import cases
case = cases.plain_cases()['nuts_iso128']
d = np.load('tests/golden/nuts_iso128.npz')
kw = dict(case['kw']); kw.pop('nuts')
for tuning in (0,):
    init = torch.from_numpy(d['init_0'])[None]
    z = torch.from_numpy(d['z_0'])[:, None]
    logu = torch.from_numpy(d['logu_0'])[:, None]
    for S in (1, 2, 3, 8):
        res = engine.hmc_run(case['target'], init, kw['num_samples'], 10, 0.1, burn=40, nuts=True, normals=z, log_uniforms=logu, record_ham=True)
        torch.cuda.synchronize()
        break
    h = res.ham[0].cpu().numpy(); a = res.accepted[0].cpu().numpy()
    for n in range(10):
        print(n, h[n], (d['ham_old_0'][n], d['ham_new_0'][n]), a[n], d['accepted_0'][n])
    print('eps', float(res.step_size[0]), d['final_step_size_0'])
