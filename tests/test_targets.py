"""CPU: every descriptor is a valid reference log_prob_func and its analytic gradient (the op order the CUDA
kernels use) equals autograd's BIT FOR BIT -- the premise of trajectory-level parity."""
import torch

from hamiltorch_b200 import targets as T


def _autograd(tgt, x):
    x = x.detach().requires_grad_()
    return torch.autograd.grad(tgt(x), x)[0]


def test_gaussian_iso_grad_bitexact():
    torch.manual_seed(0)
    for D in (3, 17, 1024):
        for normalized in (False, True):
            tgt = T.GaussianIso(D, normalized=normalized)
            x = torch.randn(D) * 3
            assert tgt(x).dim() == 0
            assert torch.equal(_autograd(tgt, x), tgt.grad(x))


def test_gaussian_diag_grad_bitexact():
    torch.manual_seed(1)
    for D in (2, 3, 48, 1000):
        mean = torch.randn(D)
        var = 0.1 + 3 * torch.rand(D)
        tgt = T.GaussianDiag(mean, var)
        for _ in range(3):
            x = torch.randn(D) * 2
            assert torch.equal(_autograd(tgt, x), tgt.grad(x))


def test_gaussian_diag_matches_torch_distribution():
    mean = torch.tensor([0.3, -1.0, 2.0])
    std = torch.tensor([.5, 1., 2.])
    tgt = T.GaussianDiag(mean, std ** 2)
    x = torch.tensor([0.1, 0.2, -0.7])
    ref = torch.distributions.MultivariateNormal(mean, torch.diag(std ** 2)).log_prob(x)
    assert abs(float(tgt(x)) - float(ref)) < 1e-5


def test_gaussian_full_grad():
    torch.manual_seed(2)
    D = 6
    A = torch.randn(D, D, dtype=torch.float64)
    cov = A @ A.t() + D * torch.eye(D, dtype=torch.float64)
    mean = torch.randn(D)
    tgt = T.GaussianFull(mean, cov=cov)
    x = torch.randn(D)
    torch.testing.assert_close(_autograd(tgt, x), tgt.grad(x), rtol=1e-5, atol=1e-6)
    ref = torch.distributions.MultivariateNormal(mean.double(), cov).log_prob(x.double())
    assert abs(float(tgt(x)) - float(ref)) < 1e-4


def test_funnel_matches_notebook_definition():
    """notebooks/hamiltorch_log_prob_examples.ipynb cell 22 (with validation off)."""
    torch.distributions.Distribution.set_default_validate_args(False)
    D = 10

    def funnel_ll(w):
        v_dist = torch.distributions.Normal(0, 3)
        ll = v_dist.log_prob(w[0])
        x_dist = torch.distributions.Normal(0, torch.exp(-w[0]) ** 0.5)
        ll += x_dist.log_prob(w[1:]).sum()
        return ll

    tgt = T.Funnel(D + 1)
    torch.manual_seed(3)
    for _ in range(5):
        w = torch.randn(D + 1)
        assert abs(float(tgt(w)) - float(funnel_ll(w))) < 1e-4 * (1 + abs(float(tgt(w))))
        torch.testing.assert_close(_autograd(tgt, w), tgt.grad(w), rtol=1e-5, atol=1e-5)


def test_const_metric_matches_the_oracle_fisher():
    """engine.const_metric (host side of the constant-metric RMHMC path) == oracle fisher() at any point for Gaussian
    targets without jitter, for both metrics; G^-1 and chol(G) are consistent with it."""
    from hamiltorch_b200 import engine
    from oracle import rmhmc_oracle as R
    D = 12
    g = torch.Generator().manual_seed(5)
    A = torch.randn(D, D, generator=g, dtype=torch.float64) / D ** 0.5
    cov = A @ A.t() + 0.5 * torch.eye(D, dtype=torch.float64)
    for tgt in (T.GaussianFull(torch.randn(D, generator=g), cov=cov),
                T.GaussianDiag(torch.randn(D, generator=g), 0.3 + torch.rand(D, generator=g)), T.GaussianIso(D)):
        for softabs in (False, True):
            ginv, lower, log_det = engine.const_metric(tgt, softabs, 0.7)
            for _ in range(2):
                q = torch.randn(D, generator=g)
                fish, lam = R.fisher(q, tgt, None, 0.7, R.SOFTABS if softabs else R.HESSIAN, None)
                fish = fish.detach()
                assert torch.allclose(lower @ lower.t(), fish, rtol=1e-5, atol=1e-6)
                assert torch.allclose(ginv @ fish, torch.eye(D), atol=2e-5)
                ref_ld = float(lam.log().sum()) if softabs else float(torch.slogdet(fish)[1])
                assert abs(log_det - ref_ld) <= 1e-5 * (1 + abs(ref_ld))


def test_const_metric_device_operands_are_cached_per_target_object_and_version():
    """Repeated RMHMC runs on the same (unmodified) Gaussian descriptor reuse the factorised metric; an in-place edit of the
    precision (torch's version counter) or another softabs constant recomputes it."""
    from hamiltorch_b200 import engine
    D = 6
    g = torch.Generator().manual_seed(9)
    A = torch.randn(D, D, generator=g, dtype=torch.float64)
    tgt = T.GaussianFull(torch.zeros(D), cov=A @ A.t() + torch.eye(D, dtype=torch.float64))
    a = engine.const_metric_device(tgt, True, 0.7, 'cpu')
    b = engine.const_metric_device(tgt, True, 0.7, 'cpu')
    assert a[0] is b[0] and a[1] is b[1]
    c = engine.const_metric_device(tgt, True, 0.9, 'cpu')
    assert c[0] is not a[0]
    h = engine.const_metric_device(tgt, False, None, 'cpu')
    tgt.prec.mul_(2.0)
    d = engine.const_metric_device(tgt, True, 0.7, 'cpu')
    assert d[0] is not a[0]
    h2 = engine.const_metric_device(tgt, False, None, 'cpu')             # HESSIAN: G = P, so G^-1 halves
    assert h2[0] is not h[0] and torch.allclose(h2[0], h[0] / 2, rtol=1e-4, atol=1e-6)
