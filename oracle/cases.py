"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Shared definitions of the golden-fixture cases.

Each case = (target descriptor, sampler arguments, per-chain seeds).  oracle/gen_golden.py runs the UNMODIFIED
reference on them (build container only) and stores results under tests/golden/; the tests rebuild the same
case objects from here and compare the oracle and the CUDA path with the stored results.
"""
import torch

from hamiltorch_b200 import targets as T


def _rand_var(dim, seed):
    g = torch.Generator().manual_seed(seed)
    return 0.25 + 1.5 * torch.rand(dim, generator=g)


def plain_cases():
    """name -> dict(target, kwargs for sample(), seeds, init)."""
    cases = {}
    # BASELINE config 1: notebooks/hamiltorch_log_prob_examples.ipynb cells 6, 9 (diag Gaussian, sigma=.5,1,2)
    cases['cfg1_gauss3'] = dict(
        target=T.GaussianDiag(torch.zeros(3), torch.tensor([.5, 1., 2.]) ** 2),
        kw=dict(num_samples=400, num_steps_per_sample=5, step_size=0.3, burn=0),
        seeds=[123], init='zeros')
    # small twin of BASELINE config 2 (D=1024 in the live GPU tests; D=256 keeps the fixture small)
    cases['iso256'] = dict(
        target=T.GaussianIso(256),
        kw=dict(num_samples=30, num_steps_per_sample=10, step_size=0.05, burn=0),
        seeds=[0, 1, 2], init='randn0.1')
    # diagonal inv_mass + burn: exercises gibbs :201, drift :296, kinetic :814 and the burn bookkeeping
    var = _rand_var(48, 7)
    cases['diag48_mass'] = dict(
        target=T.GaussianDiag(torch.linspace(-1, 1, 48), var),
        kw=dict(num_samples=60, num_steps_per_sample=8, step_size=0.35, burn=10, inv_mass=var.clone()),
        seeds=[11, 12], init='randn0.1')
    # low acceptance: large step so that rejections (incl. the first-post-burn quirk) are exercised
    cases['diag16_rejects'] = dict(
        target=T.GaussianDiag(torch.zeros(16), _rand_var(16, 3)),
        kw=dict(num_samples=80, num_steps_per_sample=3, step_size=0.9, burn=5),
        seeds=[5, 6, 7], init='randn0.1')
    # small twin of BASELINE config 5: HMC_NUTS step-size adaptation
    cases['nuts_iso128'] = dict(
        target=T.GaussianIso(128),
        kw=dict(num_samples=60, num_steps_per_sample=10, step_size=0.1, burn=40, nuts=True,
                desired_accept_rate=0.8),
        seeds=[21, 22], init='randn0.1')
    # ---- coupled targets / full mass matrix: thread-per-chain kernel (matvec summation order differs from torch's,
    #      so these compare to 'rtol' instead of bit-exactly)
    g = torch.Generator().manual_seed(21)
    A = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))[0]
    cov = A @ torch.diag(torch.tensor([.25, 1., 4.], dtype=torch.float64)) @ A.t()
    # the "3D correlated Gaussian" reading of BASELINE config 1: rotated cov diag(.25,1,4)
    cases['cfg1_corr_gauss3'] = dict(
        target=T.GaussianFull(torch.zeros(3), cov=cov),
        kw=dict(num_samples=120, num_steps_per_sample=5, step_size=0.3, burn=0),
        seeds=[123], init='zeros', rtol=2e-5)
    # the notebook funnel (D=10+1) under plain HMC and under NUTS (hamiltorch_log_prob_examples.ipynb cells 24, 26)
    # (the funnel's dynamics are chaotic: fp32 round-off differences -- expf vs Sleef exp, summation order -- grow
    #  exponentially along the chain, so these fixtures are short and compared to 1e-3)
    cases['funnel11_hmc'] = dict(
        target=T.Funnel(11), kw=dict(num_samples=12, num_steps_per_sample=25, step_size=0.2, burn=0),
        seeds=[123], init='funnel', rtol=1e-3)
    cases['funnel11_nuts'] = dict(
        target=T.Funnel(11), kw=dict(num_samples=14, num_steps_per_sample=25, step_size=0.01, burn=10, nuts=True,
                                     desired_accept_rate=0.75),
        seeds=[123], init='funnel', rtol=1e-3)
    # full (2-D) inv_mass: gibbs :199, drift :294, kinetic :812
    B = torch.randn(5, 5, generator=g)
    im_full = (B @ B.t() / 5 + torch.eye(5)).contiguous()
    cases['diag5_fullmass'] = dict(
        target=T.GaussianDiag(torch.linspace(-1, 1, 5), _rand_var(5, 11)),
        kw=dict(num_samples=50, num_steps_per_sample=6, step_size=0.25, burn=5, inv_mass=im_full),
        seeds=[31, 32], init='randn0.1', rtol=2e-5)
    # ---- D > 16: the tensor-core paths (3xTF32 GEMMs over all chains vs the reference's fp32 mv / matmul): rtol 2e-4
    cases['full48_dense'] = dict(                      # dense precision, no mass: dense_step_kernel
        target=T.GaussianFull(0.3 * torch.randn(48, generator=g), cov=_spd64(48, 61)),
        kw=dict(num_samples=14, num_steps_per_sample=6, step_size=0.25, burn=2),
        seeds=[43, 44], init='randn0.1', rtol=2e-4)
    cases['full40_fullmass'] = dict(                   # dense precision AND 2-D inv_mass: dense_lin_kernel, 2L+4 GEMMs
        target=T.GaussianFull(0.3 * torch.randn(40, generator=g), cov=_spd64(40, 62)),
        kw=dict(num_samples=12, num_steps_per_sample=5, step_size=0.2, burn=3, inv_mass=_spd64(40, 63).float()),
        seeds=[41, 42], init='randn0.1', rtol=2e-4)
    cases['iso40_fullmass_nuts'] = dict(               # element-wise target, 2-D inv_mass, step-size adaptation
        target=T.GaussianIso(40),
        kw=dict(num_samples=14, num_steps_per_sample=5, step_size=0.1, burn=8, nuts=True, desired_accept_rate=0.8,
                inv_mass=_spd64(40, 64).float()),
        seeds=[45], init='randn0.1', rtol=2e-4)
    # ---- block-list inv_mass (samplers.py:188-197, :287-292, :803-809): blocks of 2 + 3 + 1 on the thread-per-chain
    #      kernel, 16 + 24 on the tensor-core path; the kernels see the block-diagonal matrix
    cases['diag6_blockmass'] = dict(
        target=T.GaussianDiag(torch.linspace(-1, 1, 6), _rand_var(6, 12)),
        kw=dict(num_samples=40, num_steps_per_sample=5, step_size=0.3, burn=4,
                inv_mass=[_spd64(2, 81).float(), _spd64(3, 82).float(), _spd64(1, 83).float()]),
        seeds=[51, 52], init='randn0.1', rtol=2e-5)
    cases['iso40_blockmass'] = dict(
        target=T.GaussianIso(40),
        kw=dict(num_samples=12, num_steps_per_sample=5, step_size=0.15, burn=2,
                inv_mass=[_spd64(16, 84).float(), _spd64(24, 85).float()]),
        seeds=[53], init='randn0.1', rtol=2e-4)
    return cases


def _spd64(dim, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(dim, dim, generator=g, dtype=torch.float64) / dim ** 0.5
    return A @ A.t() + 0.6 * torch.eye(dim, dtype=torch.float64)


def make_init(kind, dim, seed):
    """multi_chain convention (util.py:386-389): manual_seed(seed) then the prior draw, then sample()."""
    torch.manual_seed(seed)
    if kind == 'zeros':
        return torch.zeros(dim)
    if kind == 'randn0.1':
        return 0.1 * torch.randn(dim)
    if kind == 'funnel':                           # notebook: ones(D+1) with v = 0
        x = torch.ones(dim)
        x[0] = 0.
        return x
    raise ValueError(kind)


# ----------------------------------------------------------------------------------------------------------
# Bayesian-NN cases (sample_model / sample_split_model / predict_model, samplers.py:1261-1562)
# ----------------------------------------------------------------------------------------------------------
def mlp_problem(seed=0, n=48, n_in=6, hidden=16, n_out=1, depth=1, act='ReLU', task='regression'):
    """Small synthetic problem + an nn.Sequential dense stack, deterministic in ``seed``.  task: 'regression',
    'binary' (labels in {0,1}, one logit), 'multiclass' (class indices, n_out logits), 'logsoftmax' (same, the model
    ends in LogSoftmax)."""
    import torch.nn as nn
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, n_in, generator=g)
    w = torch.randn(n_in, n_out, generator=g)
    y = torch.sin(x @ w / 2) + 0.1 * torch.randn(n, n_out, generator=g)
    if task == 'binary':
        y = (y > 0).float()
    elif task in ('multiclass', 'logsoftmax'):
        y = y.argmax(1).float()
    torch.manual_seed(seed)                    # the layers' default init draws from the global generator
    layers, last = [], n_in
    for _ in range(depth):
        layers += [nn.Linear(last, hidden), getattr(nn, act)()]
        last = hidden
    layers.append(nn.Linear(last, n_out))
    if task == 'logsoftmax':
        layers.append(nn.LogSoftmax(dim=1))
    model = nn.Sequential(*layers)
    return model, x, y


def mlp_cases():
    """name -> dict(problem kwargs, sampler kwargs).  scheme: None (sample_model, full batch) or the splitting
    integrator name; num_splits / batch as in sample_split_model."""
    base = dict(num_samples=14, num_steps_per_sample=5, burn=2, tau_out=50., seeds=[3, 4])
    cases = {
        'mlp_full': dict(problem=dict(seed=0), scheme=None, step_size=0.01, **base),
        'mlp_split_sym': dict(problem=dict(seed=1), scheme='SPLITTING', num_splits=3, step_size=0.01, **base),
        'mlp_split_rand': dict(problem=dict(seed=2), scheme='SPLITTING_RAND', num_splits=3, step_size=0.01, **base),
        'mlp_split_kmid': dict(problem=dict(seed=3), scheme='SPLITTING_KMID', num_splits=3, step_size=0.01, **base),
        'mlp_deep_tanh_mass': dict(problem=dict(seed=4, depth=2, act='Tanh', hidden=12), scheme='SPLITTING',
                                   num_splits=4, step_size=0.01, diag_mass=True,
                                   tau_list=[1., 2., .5, 1., 3., 1.], **base),
        # classification likelihoods (samplers.py:1170-1180)
        'mlp_binary': dict(problem=dict(seed=5, task='binary'), scheme=None, step_size=0.03,
                           model_loss='binary_class_linear_output', **dict(base, tau_out=1.)),
        'mlp_multiclass': dict(problem=dict(seed=6, task='multiclass', n_out=3), scheme=None, step_size=0.03,
                               model_loss='multi_class_linear_output', **dict(base, tau_out=1.)),
        'mlp_logsoftmax_split': dict(problem=dict(seed=7, task='logsoftmax', n_out=3), scheme='SPLITTING',
                                     num_splits=3, step_size=0.05, model_loss='multi_class_log_softmax_output',
                                     **dict(base, tau_out=20.)),
    }
    return cases


# ----------------------------------------------------------------------------------------------------------
# RMHMC cases (sampler=RMHMC, explicit / implicit integrators, HESSIAN / SOFTABS metrics)
# ----------------------------------------------------------------------------------------------------------
def rmhmc_cases():
    funnel_kw = dict(softabs_const=1e6, metric='SOFTABS', jitter=1e-3)
    cases = {
        # small twin of BASELINE config 3: 2-D funnel, explicit integrator, omega=10, eps=.05
        'rmhmc_exp_funnel2': dict(target=T.Funnel(2), init=[0., 1.], integrator='EXPLICIT', num_samples=12,
                                  num_steps_per_sample=5, step_size=0.05, burn=2, explicit_binding_const=10,
                                  seeds=[1, 2], **funnel_kw),
        # general-D path (the notebook funnel has D=11): 5-D
        'rmhmc_exp_funnel5': dict(target=T.Funnel(5), init=[0., 1., 1., 1., 1.], integrator='EXPLICIT', num_samples=8,
                                  num_steps_per_sample=3, step_size=0.05, burn=0, explicit_binding_const=10,
                                  seeds=[3], **funnel_kw),
        # implicit (generalised leapfrog), no jitter so that the number of metric evaluations does not touch the RNG
        'rmhmc_imp_funnel2': dict(target=T.Funnel(2), init=[0., 1.], integrator='IMPLICIT', num_samples=8,
                                  num_steps_per_sample=4, step_size=0.1, burn=1, softabs_const=1e6, metric='SOFTABS',
                                  jitter=None, fixed_point_threshold=1e-5, fixed_point_max_iterations=1000,
                                  seeds=[5, 9]),          # seeds whose chains hit no LogProbError
        # HESSIAN metric on a Gaussian (constant metric = precision)
        'rmhmc_exp_hess_gauss3': dict(target=T.GaussianDiag(torch.tensor([0., 1., -1.]),
                                                            torch.tensor([.5, 1., 2.]) ** 2),
                                      init=[0.2, 0.8, -1.5], integrator='EXPLICIT', num_samples=10,
                                      num_steps_per_sample=4, step_size=0.15, burn=0, explicit_binding_const=5,
                                      metric='HESSIAN', jitter=None, softabs_const=None, seeds=[6]),
        # ---- constant-metric tensor-core path (hmcx_rmhmc_dense_run): Gaussian targets without jitter
        'rmhmc_exp_hess_full24': dict(target=T.GaussianFull(torch.linspace(-0.5, 0.5, 24), cov=_spd64(24, 71)),
                                      init=[0.1 * ((i * 7) % 5 - 2) for i in range(24)], integrator='EXPLICIT',
                                      num_samples=6, num_steps_per_sample=3, step_size=0.6, burn=1,
                                      explicit_binding_const=10, metric='HESSIAN', jitter=None, softabs_const=None,
                                      seeds=[7, 8]),
        # Metric.JACOBIAN_DIAG (:100-106): G = diag((d log p / d theta_i)^2) + jitter
        'rmhmc_exp_jacdiag_funnel3': dict(target=T.Funnel(3), init=[0.5, 1., -1.], integrator='EXPLICIT', num_samples=8,
                                          num_steps_per_sample=3, step_size=0.05, burn=0, explicit_binding_const=10,
                                          metric='JACOBIAN_DIAG', jitter=1e-2, softabs_const=None,
                                          seeds=[10, 16]),      # seeds whose first iteration accepts: the reference
                                          # crashes (RuntimeError) after a reject at n <= burn with this metric
        'rmhmc_imp_jacdiag_gauss3': dict(target=T.GaussianDiag(torch.tensor([0., 1., -1.]),
                                                               torch.tensor([.5, 1., 2.]) ** 2),
                                         init=[0.6, 0.2, -2.5], integrator='IMPLICIT', num_samples=8,
                                         num_steps_per_sample=3, step_size=0.05, burn=0, metric='JACOBIAN_DIAG',
                                         jitter=None, softabs_const=None, fixed_point_threshold=1e-5,
                                         fixed_point_max_iterations=1000, seeds=[16]),
        # the thread-per-chain kernel on a dense precision WITH jitter (D <= 16: position-independent but per-call random metric)
        'rmhmc_exp_softabs_full5_jitter': dict(target=T.GaussianFull(torch.linspace(-0.5, 0.5, 5), cov=_spd64(5, 73)),
                                               init=[0.3, -0.2, 0.1, 0.4, -0.3], integrator='EXPLICIT', num_samples=8,
                                               num_steps_per_sample=3, step_size=0.2, burn=1,
                                               explicit_binding_const=10, metric='SOFTABS', jitter=1e-2,
                                               softabs_const=1e3, seeds=[14]),
        # ---- one-CTA-per-chain kernel (hmcx_rmhmc_cta.cu): metric assembled / eigen-decomposed in shared memory, 16 < D <= 64
        # position-dependent metric beyond the register kernel: the funnel at D = 32
        'rmhmc_exp_funnel32': dict(target=T.Funnel(32), init=[0.] + [0.5 * (((i * 5) % 7) - 3) / 3 for i in range(31)],
                                   integrator='EXPLICIT', num_samples=7, num_steps_per_sample=3, step_size=0.03, burn=1,
                                   explicit_binding_const=10, seeds=[21, 22], **funnel_kw),
        # dense precision with the reference's jitter: a per-call random metric at D = 48
        'rmhmc_exp_softabs_full48_jitter': dict(target=T.GaussianFull(torch.linspace(-0.5, 0.5, 48), cov=_spd64(48, 75)),
                                                init=[0.1 * ((i * 7) % 5 - 2) for i in range(48)], integrator='EXPLICIT',
                                                num_samples=6, num_steps_per_sample=3, step_size=0.3, burn=1,
                                                explicit_binding_const=10, metric='SOFTABS', jitter=1e-3,
                                                softabs_const=1e3, seeds=[23]),
        # implicit integrator with a position-dependent metric beyond D = 16 (the funnel's Hessian has an eigenvalue of
        # multiplicity D-2, so without jitter the reference's eigh backward is NaN: JACOBIAN_DIAG on a Gaussian instead)
        'rmhmc_imp_jacdiag_diag24': dict(target=T.GaussianDiag(torch.linspace(-1, 1, 24), _rand_var(24, 77)),
                                         init=[1.5 + 0.2 * ((i * 3) % 7) for i in range(24)], integrator='IMPLICIT',
                                         num_samples=6, num_steps_per_sample=3, step_size=0.05, burn=1,
                                         metric='JACOBIAN_DIAG', jitter=None, softabs_const=None,
                                         fixed_point_threshold=1e-5, fixed_point_max_iterations=1000, seeds=[24, 26]),
        # HESSIAN metric + jitter on a diagonal Gaussian at D = 64 (the kernel's largest size)
        'rmhmc_exp_hess_diag64_jitter': dict(target=T.GaussianDiag(torch.linspace(-1, 1, 64), _rand_var(64, 76)),
                                             init=[0.2 * ((i * 3) % 7 - 3) for i in range(64)], integrator='EXPLICIT',
                                             num_samples=5, num_steps_per_sample=2, step_size=0.2, burn=0,
                                             explicit_binding_const=10, metric='HESSIAN', jitter=1e-3,
                                             softabs_const=None, seeds=[25]),
        'rmhmc_imp_softabs_diag20': dict(target=T.GaussianDiag(torch.linspace(-1, 1, 20), _rand_var(20, 72)),
                                         init=[0.2 * ((i * 3) % 7 - 3) for i in range(20)], integrator='IMPLICIT',
                                         num_samples=6, num_steps_per_sample=3, step_size=0.5, burn=1,
                                         softabs_const=1.0, metric='SOFTABS', jitter=None, fixed_point_threshold=1e-5,
                                         fixed_point_max_iterations=1000, seeds=[9]),
    }
    return cases


def standalone_rm_cases():
    """Stand-alone samplers.leapfrog / samplers.hamiltonian with sampler=RMHMC (:305-462, :817-829) -- name -> kwargs."""
    return {
        'exp_funnel5': dict(target=T.Funnel(5), q=[0.3, 1., -1., 0.5, 0.8], integrator='EXPLICIT', metric='SOFTABS',
                            steps=3, step_size=0.05, jitter=1e-3, softabs_const=1e6, explicit_binding_const=10, seed=31),
        'exp_funnel32': dict(target=T.Funnel(32), q=[0.2] + [0.5 * (((i * 5) % 7) - 3) / 3 for i in range(31)],
                             integrator='EXPLICIT', metric='SOFTABS', steps=2, step_size=0.03, jitter=1e-3,
                             softabs_const=1e6, explicit_binding_const=10, seed=32),
        'imp_jacdiag_gauss3': dict(target=T.GaussianDiag(torch.tensor([0., 1., -1.]), torch.tensor([.5, 1., 2.]) ** 2),
                                   q=[0.6, 0.2, -2.5], integrator='IMPLICIT', metric='JACOBIAN_DIAG', steps=3,
                                   step_size=0.05, jitter=None, softabs_const=None, seed=33),
        'imp_hess_full24': dict(target=T.GaussianFull(torch.linspace(-0.5, 0.5, 24), cov=_spd64(24, 71)),
                                q=[0.1 * ((i * 7) % 5 - 2) for i in range(24)], integrator='IMPLICIT', metric='HESSIAN',
                                steps=3, step_size=0.3, jitter=None, softabs_const=None, seed=34),
    }
