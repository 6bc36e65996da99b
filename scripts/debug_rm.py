import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hamiltorch_b200 import engine
from oracle import cases
name='rmhmc_imp_funnel2'
case = cases.rmhmc_cases()[name]
d = np.load('tests/golden/%s.npz' % name)
S, L = case['num_samples'], case['num_steps_per_sample']
init = torch.tensor(case['init']).repeat(1, 1)
z = torch.from_numpy(d['z_0'])[:, None]; logu = torch.from_numpy(d['logu_0'])[:, None]
res = engine.rmhmc_run(case['target'], init, S, L, case['step_size'], burn=case['burn'], jitter=None, softabs_const=1e6,
                       fixed_point_threshold=1e-5, fixed_point_max_iterations=1000, explicit=False, softabs=True,
                       normals=z, log_uniforms=logu, record_ham=True)
torch.cuda.synchronize()
h = res.ham[0].cpu().numpy(); a = res.accepted[0].cpu().numpy()
for n in range(S):
    print(n, h[n], (d['ham_old_0'][n], d['ham_new_0'][n]), a[n], d['accepted_0'][n], d['logu_0'][n])
print(res.samples[0].cpu().numpy()); print(d['samples_0'])
