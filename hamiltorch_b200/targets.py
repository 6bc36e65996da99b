"""Native target descriptors.

The reference takes an opaque Python ``log_prob_func`` (samplers.py:857-858) and differentiates it with
autograd (samplers.py:65).  An opaque callable cannot enter a CUDA kernel, so the B200 engine recognises a
small family of *descriptors*.  Every descriptor is

* a valid reference ``log_prob_func`` (``__call__`` takes a 1-D tensor, returns a scalar, is written with plain
  torch ops) -- the *same object* drives the oracle / the unmodified reference and the CUDA kernels; and
* a description of the analytic gradient the kernels evaluate, with the fp32 operation order chosen so that
  ``torch.autograd.grad(self(x), x)`` and the in-kernel gradient agree bit for bit (``grad`` restates that
  order in torch ops; tests/test_targets.py pins the equality on CPU).

Anything else passed as ``log_prob_func`` is refused by ``hamiltorch_b200.sample`` with a clear error: there
is no CPU fallback on the product path.
"""
import math

import torch

# target kinds -- must match include/hmcx.h
KIND_GAUSS_ISO = 0
KIND_GAUSS_DIAG = 1
KIND_GAUSS_FULL = 2
KIND_FUNNEL = 3
KIND_MLP = 4

_LOG_2PI = math.log(2.0 * math.pi)


class Target:
    """Base class: a log-density the sm_100a kernels know how to differentiate."""

    kind = -1
    dim = 0

    def __call__(self, x):  # pragma: no cover - abstract
        raise NotImplementedError

    def grad(self, x):  # pragma: no cover - abstract
        raise NotImplementedError

    def _tensors(self):
        return {}

    def to(self, device):
        """Return a copy of the descriptor whose parameter tensors live on ``device``."""
        import copy
        new = copy.copy(self)
        for name, t in self._tensors().items():
            setattr(new, name, t.to(device))
        return new


class GaussianIso(Target):
    """log p(x) = -0.5 * sum(x*x)  (+ -0.5*D*log(2*pi) when ``normalized``).

    BASELINE config 2 / 5 target.  Gradient: -x (exact in fp32).
    """

    kind = KIND_GAUSS_ISO

    def __init__(self, dim, normalized=False):
        self.dim = int(dim)
        self.log_norm = float(-0.5 * self.dim * _LOG_2PI) if normalized else 0.0

    def __call__(self, x):
        lp = -0.5 * (x * x).sum()
        if self.log_norm != 0.0:
            lp = lp + self.log_norm
        return lp

    def grad(self, x):
        return -x


class GaussianDiag(Target):
    """Independent Gaussian, log p(x) = -0.5 * sum((x-mean)^2 * inv_var) + log_norm.

    BASELINE config 1 (the notebook's diag-covariance 3-D Gaussian, notebooks/hamiltorch_log_prob_examples.ipynb
    cell 6) and the reference's own reversibility test target (tests/test_util.py:98-101).
    fp32 op order (shared with the kernel): y = x-mean; g = -(inv_var*y).
    """

    kind = KIND_GAUSS_DIAG

    def __init__(self, mean, var, normalized=True):
        mean = torch.as_tensor(mean, dtype=torch.float32).flatten().clone()
        var = torch.as_tensor(var, dtype=torch.float32).flatten().clone()
        if mean.shape != var.shape:
            raise ValueError('mean and var must have the same length')
        self.dim = mean.numel()
        self.mean = mean
        self.inv_var = 1.0 / var
        if normalized:
            self.log_norm = float(-0.5 * (self.dim * _LOG_2PI + torch.log(var.double()).sum().item()))
        else:
            self.log_norm = 0.0

    def _tensors(self):
        return {'mean': self.mean, 'inv_var': self.inv_var}

    def __call__(self, x):
        y = x - self.mean
        lp = -0.5 * ((y * y) * self.inv_var).sum()
        if self.log_norm != 0.0:
            lp = lp + self.log_norm
        return lp

    def grad(self, x):
        return -(self.inv_var * (x - self.mean))


class GaussianFull(Target):
    """Correlated Gaussian with precision matrix P: log p(x) = -0.5 * y.(P y) + log_norm, y = x-mean.

    P is symmetrised at construction.  Gradient (kernel order): g = -(P y) with the row dot products
    accumulated in fp32.
    """

    kind = KIND_GAUSS_FULL

    def __init__(self, mean, cov=None, prec=None, normalized=True):
        mean = torch.as_tensor(mean, dtype=torch.float32).flatten().clone()
        if (cov is None) == (prec is None):
            raise ValueError('give exactly one of cov / prec')
        if prec is None:
            prec = torch.linalg.inv(torch.as_tensor(cov, dtype=torch.float64))
        prec = torch.as_tensor(prec, dtype=torch.float64)
        prec = 0.5 * (prec + prec.t())
        self.dim = mean.numel()
        if prec.shape != (self.dim, self.dim):
            raise ValueError('precision must be (D, D)')
        self.mean = mean
        self.prec = prec.to(torch.float32).contiguous()
        if normalized:
            self.log_norm = float(-0.5 * (self.dim * _LOG_2PI - torch.linalg.slogdet(prec)[1].item()))
        else:
            self.log_norm = 0.0

    def _tensors(self):
        return {'mean': self.mean, 'prec': self.prec}

    def __call__(self, x):
        y = x - self.mean
        lp = -0.5 * torch.dot(y, torch.mv(self.prec, y))
        if self.log_norm != 0.0:
            lp = lp + self.log_norm
        return lp

    def grad(self, x):
        return -torch.mv(self.prec, x - self.mean)


class Funnel(Target):
    """Neal's funnel as in notebooks/hamiltorch_log_prob_examples.ipynb cell 22:
    v = w[0] ~ N(0, sigma_v^2),  w[1:] ~ N(0, exp(-v)).

    Closed form (SURVEY.md section 8d cfg 3; avoids torch.distributions' scale validation, section 8a quirks):
      log p = -v^2/(2 sigma_v^2) - 0.5*log(2 pi sigma_v^2) + n*(0.5*v - 0.5*log 2pi) - 0.5*exp(v)*sum(x^2)
    """

    kind = KIND_FUNNEL

    def __init__(self, dim, sigma_v=3.0):
        if dim < 2:
            raise ValueError('funnel needs dim >= 2')
        self.dim = int(dim)
        self.sigma_v = float(sigma_v)
        self.inv_var_v = 1.0 / (self.sigma_v ** 2)
        n = self.dim - 1
        self.log_norm = float(-0.5 * math.log(2.0 * math.pi * self.sigma_v ** 2) - 0.5 * n * _LOG_2PI)

    def __call__(self, w):
        v = w[0]
        x = w[1:]
        n = self.dim - 1
        return (-0.5 * self.inv_var_v) * (v * v) + (0.5 * n) * v - 0.5 * torch.exp(v) * (x * x).sum() + self.log_norm

    def grad(self, w):
        v = w[0]
        x = w[1:]
        n = self.dim - 1
        ev = torch.exp(v)
        gv = -(self.inv_var_v * v) + 0.5 * n - 0.5 * ev * (x * x).sum()
        gx = -(ev * x)
        return torch.cat([gv.reshape(1), gx])


def is_target(obj):
    return isinstance(obj, Target)
