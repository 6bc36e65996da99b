"""Host-side mirror of the reference's sampler surface (hamiltorch/samplers.py) on top of the sm_100a kernels.

Same names, argument order, defaults and error behaviour as the reference for the hot path:
``sample`` (samplers.py:850), ``leapfrog`` (:205), ``hamiltonian`` (:738), ``gibbs`` (:152), ``acceptance``
(:609), ``adaptation`` (:629) and the ``Sampler`` / ``Integrator`` / ``Metric`` enums (:11-31).

What differs by design
  * ``log_prob_func`` must be a target descriptor from ``hamiltorch_b200.targets`` (or the list of split
    descriptors ``define_split_model_log_prob`` builds): an opaque Python callable cannot enter a CUDA kernel
    and there is no CPU fallback -> ``TypeError``.
  * the work runs on the current CUDA device whatever ``params_init.device`` is; results are returned on
    ``params_init.device`` so CPU-tensor user code keeps working.
  * extra keyword-only arguments (``rng``, ``seed``) select the random stream; ``sample_chains`` is the batched
    many-chains entry (the engine's native shape), ``sample`` is the one-chain drop-in.
"""
from enum import Enum

import torch

from . import _native as N
from . import engine
from . import targets as T
from . import util


class Sampler(Enum):          # samplers.py:11-14
    HMC = 1
    RMHMC = 2
    HMC_NUTS = 3


class Integrator(Enum):       # samplers.py:19-25
    EXPLICIT = 1
    IMPLICIT = 2
    S3 = 3
    SPLITTING = 4
    SPLITTING_RAND = 5
    SPLITTING_KMID = 6


class Metric(Enum):           # samplers.py:28-31
    HESSIAN = 1
    SOFTABS = 2
    JACOBIAN_DIAG = 3


_SPLIT_INTEGRATORS = (Integrator.SPLITTING, Integrator.SPLITTING_RAND, Integrator.SPLITTING_KMID)


def _require_target(log_prob_func):
    if isinstance(log_prob_func, list):
        if not all(T.is_target(f) for f in log_prob_func):
            raise TypeError('every element of a split log_prob_func list must be a hamiltorch_b200 target')
        return
    if not T.is_target(log_prob_func):
        raise TypeError(
            'hamiltorch_b200 runs the leapfrog loop inside a CUDA kernel and cannot call an opaque Python '
            'log_prob_func.  Pass a descriptor from hamiltorch_b200.targets (GaussianIso, GaussianDiag, '
            'GaussianFull, Funnel) or use sample_model / sample_split_model for Bayesian NNs.  '
            'There is no CPU fallback.')


# ----------------------------------------------------------------------------------------------------------
# small pieces of the reference surface
# ----------------------------------------------------------------------------------------------------------
def acceptance(h_old, h_new):
    """samplers.py:609-626: log acceptance ratio as a Python float."""
    return float(-h_new + h_old)


def adaptation(rho, t, step_size_init, H_t, eps_bar, desired_accept_rate=0.8):
    """samplers.py:629-674, host restatement used only for API completeness (the in-kernel version is what the
    sampler runs).  Same mixed double / fp32 arithmetic."""
    t = t + 1
    if util.has_nan_or_inf(torch.tensor([rho])):
        alpha = 0
    else:
        alpha = min(1., float(torch.exp(torch.FloatTensor([rho]))))
    mu = float(torch.log(10 * torch.FloatTensor([step_size_init])))
    w = 1 / (t + 10)
    H_t = (1 - w) * H_t + w * (desired_accept_rate - alpha)
    x_new = mu - (t ** 0.5) / 0.05 * H_t
    step_size = float(torch.exp(torch.FloatTensor([x_new])))
    x_new_bar = t ** -0.75 * x_new + (1 - t ** -0.75) * torch.log(torch.FloatTensor([eps_bar]))
    eps_bar = float(torch.exp(x_new_bar))
    return step_size, eps_bar, H_t


def gibbs(params, sampler=Sampler.HMC, log_prob_func=None, jitter=None, normalizing_const=1., softabs_const=None,
          mass=None, metric=Metric.HESSIAN):
    """samplers.py:152-202 -- momentum refresh.  Drawn from torch's generator in the reference's way (this is a
    host-side convenience; inside ``sample`` the draw happens in the kernel or is pre-drawn in bulk)."""
    if sampler == Sampler.RMHMC:
        raise NotImplementedError('RMHMC momentum refresh happens inside the RMHMC kernel')
    if mass is None:
        return torch.randn(params.shape, dtype=params.dtype, device=params.device)
    if isinstance(mass, list):                                          # :188-197
        samples = torch.zeros_like(params)
        i = 0
        for block in mass:
            it = block[0].shape[0]
            samples[i:it + i] = torch.distributions.MultivariateNormal(torch.zeros_like(block[0]), block).sample()
            i += it
        return samples
    if mass.dim() == 2:
        return torch.distributions.MultivariateNormal(torch.zeros_like(params), mass).sample()
    return torch.normal(torch.zeros_like(params), mass ** 0.5)


def leapfrog(params, momentum, log_prob_func, steps=10, step_size=0.1, jitter=0.01, normalizing_const=1.,
             softabs_const=1e6, explicit_binding_const=100, fixed_point_threshold=1e-20,
             fixed_point_max_iterations=6, jitter_max_tries=10, inv_mass=None, ham_func=None, sampler=Sampler.HMC,
             integrator=Integrator.IMPLICIT, metric=Metric.HESSIAN, store_on_GPU=True, debug=False, pass_grad=None,
             *, rng_uniforms=None, rng_perms=None):
    """samplers.py:205-606: plain HMC (:269-304), the SPLITTING integrators on a list of data-split closures (:465-603)
    and sampler=RMHMC (:305-462) on the GPU; returns ``(ret_params, ret_momenta)``: two lists of ``steps`` tensors (plain
    HMC: the last momentum carries the half-step correction, :302).  ``params`` may be (D,) as in the reference or (C, D)
    for C chains at once.  ``rng_uniforms`` / ``rng_perms`` inject the jitter draws (RMHMC) / the randperm of
    SPLITTING_RAND instead of consuming torch's generator."""
    _require_target(log_prob_func)
    if sampler == Sampler.HMC and integrator not in _SPLIT_INTEGRATORS:
        if pass_grad is not None:
            raise NotImplementedError('pass_grad: gradients are analytic inside the kernel')
        q_traj, p_traj = engine.leapfrog(log_prob_func, params, momentum, steps, step_size, inv_mass=inv_mass,
                                         return_trajectory=True)
        if params.dim() == 1:
            q_traj, p_traj = q_traj[:, 0], p_traj[:, 0]
        dev = params.device
        return [t.to(dev) for t in q_traj.unbind(0)], [t.to(dev) for t in p_traj.unbind(0)]
    if sampler == Sampler.HMC:
        if type(log_prob_func) is not list:
            raise RuntimeError('For splitting log_prob_func must be list of functions')       # :466-467
        if pass_grad is not None:
            raise RuntimeError('Passing user-determined gradients not implemented for splitting')
        M = len(log_prob_func)
        if M == 1 and integrator in (Integrator.SPLITTING, Integrator.SPLITTING_KMID):
            raise RuntimeError('For symmetric splitting log_prob_func must be list of functions greater than '
                               'length 1')                                                      # :497-498, :577-578
        if integrator not in (Integrator.SPLITTING, Integrator.SPLITTING_RAND, Integrator.SPLITTING_KMID):
            raise NotImplementedError()
        scheme = {Integrator.SPLITTING: N.SCHEME_SPLIT_SYM, Integrator.SPLITTING_RAND: N.SCHEME_SPLIT_RAND,
                  Integrator.SPLITTING_KMID: N.SCHEME_SPLIT_KMID}[integrator]
        perms = None
        if integrator == Integrator.SPLITTING_RAND:
            C_ = 1 if params.dim() == 1 else params.shape[0]
            if rng_perms is not None:                           # (C, M) injected permutations
                perms = rng_perms
            else:                                               # the reference's draw: torch.randperm(M) per call (:550)
                perms = torch.stack([torch.randperm(M) for _ in range(C_)])
        q_traj, p_traj = engine.split_leapfrog(log_prob_func, params, momentum, steps, step_size, scheme,
                                               inv_mass=inv_mass, perms=perms)
        if params.dim() == 1:
            q_traj, p_traj = q_traj[:, 0], p_traj[:, 0]
        dev = params.device
        return [t.to(dev) for t in q_traj.unbind(0)], [t.to(dev) for t in p_traj.unbind(0)]
    if sampler == Sampler.RMHMC:
        if pass_grad is not None:
            raise RuntimeError('Passing user-determined gradients not implemented for RMHMC')  # :310, :390-391
        if integrator not in (Integrator.EXPLICIT, Integrator.IMPLICIT):
            raise NotImplementedError()                                                         # S3: :606
        _check_rmhmc_target(log_prob_func, metric)
        if jitter is not None and rng_uniforms is None and not torch.is_tensor(jitter):
            # the reference draws torch.rand(D) inside every fisher() call (:115): a data-dependent number of draws for
            # the implicit integrator and for NaN retries, so they are drawn on the fly inside the kernel (Philox keyed
            # by a seed taken from torch's generator) unless the caller injects them
            pass
        explicit = integrator == Integrator.EXPLICIT
        seed = int(torch.randint(0, 2 ** 62, (1,)))
        qt, ptj, qcopy, pcopy, failed = engine.rmhmc_leapfrog(
            log_prob_func, params, momentum, steps, step_size, jitter=jitter, softabs_const=softabs_const,
            explicit_binding_const=explicit_binding_const, fixed_point_threshold=fixed_point_threshold,
            fixed_point_max_iterations=fixed_point_max_iterations, jitter_max_tries=jitter_max_tries, explicit=explicit,
            softabs=(metric == Metric.SOFTABS), jacdiag=(metric == Metric.JACOBIAN_DIAG), uniforms=rng_uniforms, seed=seed)
        if int(failed.sum()) > 0:
            raise util.LogProbError()                                                           # :110-112, :717, :733
        dev = params.device
        single = params.dim() == 1
        ret_params = [(t[0] if single else t).to(dev) for t in qt.unbind(0)]
        ret_momenta = [(t[0] if single else t).to(dev) for t in ptj.unbind(0)]
        if explicit:                                                                            # :462
            return [ret_params, (qcopy[0] if single else qcopy).to(dev)], \
                   [ret_momenta, (pcopy[0] if single else pcopy).to(dev)]
        return ret_params, ret_momenta
    raise NotImplementedError()


def _check_rmhmc_target(log_prob_func, metric):
    if isinstance(log_prob_func, list) or not isinstance(log_prob_func, (T.Funnel, T.GaussianIso, T.GaussianDiag,
                                                                         T.GaussianFull)):
        raise NotImplementedError('RMHMC needs closed-form third derivatives: Funnel / Gaussian descriptors')
    if metric not in (Metric.HESSIAN, Metric.SOFTABS, Metric.JACOBIAN_DIAG):
        raise NotImplementedError()
    if log_prob_func.dim > 64:
        raise NotImplementedError('stand-alone RMHMC leapfrog / hamiltonian: D <= 64 (metric in shared memory)')


def hamiltonian(params, momentum, log_prob_func, jitter=0.01, normalizing_const=1., softabs_const=1e6,
                explicit_binding_const=100, inv_mass=None, ham_func=None, sampler=Sampler.HMC,
                integrator=Integrator.EXPLICIT, metric=Metric.HESSIAN, *, rng_uniforms=None):
    """samplers.py:738-846: sampler=HMC (:779-815) and sampler=RMHMC (:817-829 -> rm_hamiltonian).  Raises util.LogProbError on a non-finite log-prob like the
    reference (:783-785).  (D,) -> tensor of shape (); (C, D) -> (C,)."""
    _require_target(log_prob_func)
    if sampler == Sampler.RMHMC:                                                                # :817-829
        _check_rmhmc_target(log_prob_func, metric)
        H, flags = engine.rmhmc_hamiltonian(log_prob_func, params, momentum, jitter=jitter, softabs_const=softabs_const,
                                            softabs=(metric == Metric.SOFTABS), jacdiag=(metric == Metric.JACOBIAN_DIAG),
                                            uniforms=rng_uniforms, seed=int(torch.randint(0, 2 ** 62, (1,))))
        if int(flags.sum()) > 0:
            raise util.LogProbError()
        if integrator == Integrator.EXPLICIT:
            H = 2 * H                                                                           # :822
        H = H.to(params.device)
        return H[0].reshape(1, 1) if params.dim() == 1 else H       # the reference returns a (1, 1) tensor (:731)
    if sampler != Sampler.HMC or isinstance(log_prob_func, list):
        raise NotImplementedError()
    H, flags = engine.hamiltonian(log_prob_func, params, momentum, inv_mass=inv_mass)
    if int(flags.sum()) > 0:
        raise util.LogProbError()
    H = H.to(params.device)
    return H[0] if params.dim() == 1 else H


# ----------------------------------------------------------------------------------------------------------
# sample()
# ----------------------------------------------------------------------------------------------------------
def _check_sample_args(params_init_dim_ok, num_samples, burn, sampler):
    if not params_init_dim_ok:
        raise RuntimeError('params_init must be a 1d tensor.')                  # :925-926
    if burn >= num_samples:
        raise RuntimeError('burn must be less than num_samples.')               # :928-929
    if sampler == Sampler.HMC_NUTS and burn == 0:
        raise RuntimeError('burn must be greater than 0 for NUTS.')             # :933-934


def _draw_reference_stream(dim, num_samples, device, num_perm=0, blocks=None):
    """Pre-draw one chain's randoms from torch's GLOBAL generators in exactly the order the reference consumes
    them: per iteration the momentum normals -- ``Normal(zeros_like(params),
    ones_like(params)).sample()`` (:186, :202), i.e. the generator of params' device -- then ``torch.rand(1)`` on
    the CPU generator (:1004).  Valid while no LogProbError occurs (that skips the iteration's rand(1))."""
    z = torch.empty((num_samples, dim), dtype=torch.float32, device=device)
    logu = torch.empty(num_samples, dtype=torch.float32)
    perms = torch.empty((num_samples, num_perm), dtype=torch.int32) if num_perm else None
    for n in range(num_samples):
        if blocks:                                          # block-list mass: one draw per block (:188-197)
            z[n] = torch.cat([torch.randn(b, dtype=torch.float32, device=device) for b in blocks])
        else:
            z[n] = torch.randn(dim, dtype=torch.float32, device=device)
        if num_perm:
            perms[n] = torch.randperm(num_perm)             # SPLITTING_RAND: once per trajectory (:550)
        logu[n] = torch.log(torch.rand(1))[0]
    if num_perm:
        return z, logu, perms
    return z, logu


def sample(log_prob_func, params_init, num_samples=10, num_steps_per_sample=10, step_size=0.1, burn=0, jitter=None,
           inv_mass=None, normalizing_const=1., softabs_const=None, explicit_binding_const=100,
           fixed_point_threshold=1e-5, fixed_point_max_iterations=1000, jitter_max_tries=10, sampler=Sampler.HMC,
           integrator=Integrator.IMPLICIT, metric=Metric.HESSIAN, debug=False, desired_accept_rate=0.8,
           store_on_GPU=True, pass_grad=None, verbose=True, *, rng='reference', seed=None):
    """Drop-in for ``hamiltorch.sample`` (samplers.py:850-1091): ONE chain, same arguments, same return value --
    a list of ``num_samples - burn`` detached (D,) tensors whose element 0 is ``params_init`` (:959), plus the
    adapted step size (NUTS) or the acceptance rate when ``debug == 2`` (:1086-1089).

    rng='reference' (default): the chain consumes torch's global random stream exactly like the reference, so
        after ``set_random_seed(s)`` the returned samples equal the reference's (to fp32 summation order in the
        Hamiltonian).  rng='philox': in-kernel counter RNG keyed by ``seed`` (default: drawn from torch's RNG).
    """
    _check_sample_args(params_init.dim() == 1, num_samples, burn, sampler)
    _require_target(log_prob_func)
    if pass_grad is not None:
        if sampler == Sampler.RMHMC:
            raise RuntimeError('Passing user-determined gradients not implemented for RMHMC')     # :310
        if integrator in _SPLIT_INTEGRATORS:
            raise RuntimeError('Passing user-determined gradients not implemented for splitting')  # :468-469
        raise NotImplementedError('pass_grad: gradients are analytic inside the kernel')
    res = _run_chains(log_prob_func, params_init.unsqueeze(0), num_samples, num_steps_per_sample, step_size, burn,
                      jitter, inv_mass, softabs_const, explicit_binding_const, fixed_point_threshold,
                      fixed_point_max_iterations, jitter_max_tries, sampler, integrator, metric,
                      desired_accept_rate, rng=rng, seed=seed, record_ham=(debug == 1),
                      sink=dict(host_samples=True) if (
                          not store_on_GPU and sampler in (Sampler.HMC, Sampler.HMC_NUTS) and
                          isinstance(log_prob_func, (T.GaussianIso, T.GaussianDiag)) and
                          integrator not in _SPLIT_INTEGRATORS and
                          not (torch.is_tensor(inv_mass) and inv_mass.dim() == 2)) else None)
    if not res.samples_padded.is_cuda:
        torch.cuda.current_stream().synchronize()       # the kernel wrote the samples into pinned host memory
    nuts = sampler == Sampler.HMC_NUTS
    out_dev = params_init.device if store_on_GPU else torch.device('cpu')
    samples = res.samples[0].to(out_dev)
    ret = list(samples.unbind(0))
    num_rejected = int(res.num_rejected[0])
    final_eps = float(res.step_size[0])
    if debug == 1:
        ham = res.ham[0].cpu()
        acc = res.accepted[0].cpu()
        for n in range(num_samples):
            print('Step: {}, Current Hamiltoninian: {}, Proposed Hamiltoninian: {}'.format(n, ham[n, 0], ham[n, 1]))
            print('Accept rho: {}'.format(min(0., float(ham[n, 0] - ham[n, 1]))) if acc[n] else 'REJECT')
    if nuts:
        print('Final Adapted Step Size: ', final_eps)                                              # :1035
    if verbose:
        print('Acceptance Rate {:.2f}'.format(1 - num_rejected / num_samples))                     # :1085
    if nuts and debug == 2:
        return ret, final_eps
    elif debug == 2:
        return ret, 1 - num_rejected / num_samples
    return ret


def sample_chains(log_prob_func, params_init, num_samples=10, num_steps_per_sample=10, step_size=0.1, burn=0,
                  jitter=None, inv_mass=None, softabs_const=None, explicit_binding_const=100,
                  fixed_point_threshold=1e-5, fixed_point_max_iterations=1000, jitter_max_tries=10,
                  sampler=Sampler.HMC, integrator=Integrator.IMPLICIT, metric=Metric.HESSIAN,
                  desired_accept_rate=0.8, rng='philox', seed=0, chain_offset=0, normals=None, log_uniforms=None,
                  record_ham=False, out=None, perms=None, uniforms=None, thin=1, moments=False, keep_samples=True,
                  store_on_GPU=True, host_windows=0):
    """The engine's native entry: C independent chains at once.  ``params_init`` is (C, D); every chain gets the
    reference's ``sample`` semantics.  Returns an ``engine.HMCResult`` whose ``.samples`` is (C, S-burn, D) on the
    GPU (row c = what ``sample`` would have returned for chain c, stacked).

    rng='philox'  in-kernel Philox4x32-10 keyed by (seed, chain_offset + c, iteration) -- results do not depend on
                  how chains are sharded over GPUs.
    rng='injected'  consume ``normals`` (S, C, D) / ``log_uniforms`` (S, C) (parity mode; SPLITTING_RAND also
                  ``perms`` (S, C, M)).
    ``log_prob_func`` may be a Gaussian / funnel descriptor, an ``MLPRegression`` (sample_model) or the list of split
    descriptors ``define_split_model_log_prob`` returns (with a SPLITTING integrator).

    Sample sink (plain HMC / HMC_NUTS on GaussianIso / GaussianDiag): ``thin`` keeps every thin-th post-burn state;
    ``moments=True`` returns per-chain running sums / sums of squares over all post-burn iterations
    (``.moment_sum``, ``.moment_sumsq``, ``.moment_count``); ``keep_samples=False`` stores no samples;
    ``store_on_GPU=False`` (the reference's flag, samplers.py:1008-1012) streams the retained samples from the kernel
    straight into pinned host memory: ``.samples`` is then a CPU tensor (synchronise the stream before reading).
    ``out=<pinned host block>, host_windows=W`` (W >= 2) delivers into the caller's block through the copy engine instead:
    the run is cut into W windows of iterations and each window's sample slots go to the host on a second stream while the
    next window computes (costs a device staging block; 57 instead of ~52.5 GB/s over PCIe 5 on B200).
    """
    if params_init.dim() != 2:
        raise RuntimeError('sample_chains: params_init must be (num_chains, D)')
    _check_sample_args(True, num_samples, burn, sampler)
    _require_target(log_prob_func)
    return _run_chains(log_prob_func, params_init, num_samples, num_steps_per_sample, step_size, burn, jitter,
                       inv_mass, softabs_const, explicit_binding_const, fixed_point_threshold,
                       fixed_point_max_iterations, jitter_max_tries, sampler, integrator, metric,
                       desired_accept_rate, rng=rng, seed=seed, chain_offset=chain_offset, normals=normals,
                       log_uniforms=log_uniforms, record_ham=record_ham, out=out, injected_perms=perms,
                       injected_uniforms=uniforms,
                       sink=dict(thin=thin, moments=moments, keep_samples=keep_samples, host_samples=not store_on_GPU,
                                 host_windows=host_windows))


def _run_chains(log_prob_func, q0, num_samples, L, step_size, burn, jitter, inv_mass, softabs_const,
                explicit_binding_const, fixed_point_threshold, fixed_point_max_iterations, jitter_max_tries,
                sampler, integrator, metric, desired_accept_rate, rng='philox', seed=None, chain_offset=0,
                normals=None, log_uniforms=None, record_ham=False, out=None, injected_perms=None,
                injected_uniforms=None, sink=None):
    nuts = sampler == Sampler.HMC_NUTS
    sink = sink or {}
    if (sink.get('thin', 1) != 1 or sink.get('moments') or not sink.get('keep_samples', True) or
            sink.get('host_samples')) and not (sampler in (Sampler.HMC, Sampler.HMC_NUTS) and
                                               isinstance(log_prob_func, (T.GaussianIso, T.GaussianDiag)) and
                                               integrator not in _SPLIT_INTEGRATORS and
                                               not (torch.is_tensor(inv_mass) and inv_mass.dim() == 2) and
                                               not isinstance(inv_mass, list)):
        raise NotImplementedError('thin / moments / keep_samples / store_on_GPU=False: plain HMC on element-wise targets '
                                  'with inv_mass None or 1-D')
    if nuts:
        sampler = Sampler.HMC                                                     # :932-936
    if sampler == Sampler.HMC and integrator not in _SPLIT_INTEGRATORS:
        if isinstance(log_prob_func, list):
            raise NotImplementedError('a list log_prob_func needs a SPLITTING integrator')
        D = log_prob_func.dim
        if q0.shape[1] != D:
            raise RuntimeError('params_init has %d entries, the target has %d' % (q0.shape[1], D))
        if rng == 'reference':
            if q0.shape[0] != 1:
                raise RuntimeError("rng='reference' replays torch's global stream and is defined for one chain")
            z, logu = _draw_reference_stream(D, num_samples, q0.device,
                                             blocks=[b.shape[0] for b in inv_mass] if isinstance(inv_mass, list) else None)
            normals, log_uniforms = z.unsqueeze(1), logu.unsqueeze(1)
        elif rng == 'injected':
            if normals is None or log_uniforms is None:
                raise RuntimeError("rng='injected' needs normals and log_uniforms")
        elif rng == 'philox':
            normals = log_uniforms = None
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)))
        else:
            raise ValueError('unknown rng mode %r' % (rng,))
        return engine.hmc_run(log_prob_func, q0, num_samples, L, step_size, burn=burn, inv_mass=inv_mass, nuts=nuts,
                              desired_accept_rate=desired_accept_rate, seed=seed or 0, chain_offset=chain_offset,
                              normals=normals, log_uniforms=log_uniforms, record_ham=record_ham, out=out,
                              scheme=N.SCHEME_PLAIN if isinstance(log_prob_func, T.MLPRegression) else None, **sink)
    if sampler == Sampler.HMC:
        if type(log_prob_func) is not list:
            raise RuntimeError('For splitting log_prob_func must be list of functions')            # :466-467
        M = len(log_prob_func)
        if M == 1 and integrator in (Integrator.SPLITTING, Integrator.SPLITTING_KMID):
            raise RuntimeError('For symmetric splitting log_prob_func must be list of functions greater than '
                               'length 1')                                                          # :497-498, :577-578
        if isinstance(inv_mass, list):
            raise NotImplementedError('block-list inv_mass with a splitting integrator: the reference ignores the blocks in '
                                      'the drift (samplers.py:514-515); pass the block-diagonal matrix or a 1-D inv_mass')
        scheme = {Integrator.SPLITTING: N.SCHEME_SPLIT_SYM, Integrator.SPLITTING_RAND: N.SCHEME_SPLIT_RAND,
                  Integrator.SPLITTING_KMID: N.SCHEME_SPLIT_KMID}[integrator]
        D = log_prob_func[0].dim
        if q0.shape[1] != D:
            raise RuntimeError('params_init has %d entries, the target has %d' % (q0.shape[1], D))
        perms = None
        if rng == 'reference':
            if q0.shape[0] != 1:
                raise RuntimeError("rng='reference' replays torch's global stream and is defined for one chain")
            drawn = _draw_reference_stream(D, num_samples, q0.device,
                                           M if integrator == Integrator.SPLITTING_RAND else 0)
            z, logu, pm = drawn if len(drawn) == 3 else (drawn[0], drawn[1], None)
            normals, log_uniforms = z.unsqueeze(1), logu.unsqueeze(1)
            perms = None if pm is None else pm.unsqueeze(1)
        elif rng == 'injected':
            if normals is None or log_uniforms is None:
                raise RuntimeError("rng='injected' needs normals and log_uniforms")
            perms = injected_perms
        elif rng == 'philox':
            normals = log_uniforms = None
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)))
        else:
            raise ValueError('unknown rng mode %r' % (rng,))
        return engine.hmc_run(log_prob_func, q0, num_samples, L, step_size, burn=burn, inv_mass=inv_mass, nuts=nuts,
                              desired_accept_rate=desired_accept_rate, seed=seed or 0, chain_offset=chain_offset,
                              normals=normals, log_uniforms=log_uniforms, record_ham=record_ham, out=out,
                              scheme=scheme, perms=perms)
    if sampler == Sampler.RMHMC and integrator in (Integrator.EXPLICIT, Integrator.IMPLICIT):
        if isinstance(log_prob_func, list) or not isinstance(log_prob_func, (T.Funnel, T.GaussianIso, T.GaussianDiag,
                                                                             T.GaussianFull)):
            raise NotImplementedError('RMHMC needs closed-form third derivatives: Funnel / Gaussian descriptors')
        jacdiag = metric == Metric.JACOBIAN_DIAG
        if metric not in (Metric.HESSIAN, Metric.SOFTABS, Metric.JACOBIAN_DIAG):
            raise NotImplementedError()
        Dd = log_prob_func.dim
        const_metric = jitter is None and not jacdiag and isinstance(log_prob_func, (T.GaussianIso, T.GaussianDiag,
                                                                                     T.GaussianFull))
        if Dd > 64 and not const_metric:
            raise NotImplementedError('RMHMC with a position-dependent or jittered metric: D <= 64 (metric, eigenvectors '
                                      'and one work matrix in shared memory); Gaussian targets with jitter=None run on '
                                      'the constant-metric tensor-core path at any D')
        if inv_mass is not None:
            pass                                            # the reference ignores inv_mass for RMHMC (:989 comment)
        D = log_prob_func.dim
        if q0.shape[1] != D:
            raise RuntimeError('params_init has %d entries, the target has %d' % (q0.shape[1], D))
        explicit = integrator == Integrator.EXPLICIT
        uniforms = None
        if rng == 'reference':
            if q0.shape[0] != 1:
                raise RuntimeError("rng='reference' replays torch's global stream and is defined for one chain")
            if jitter is not None and not explicit:
                raise NotImplementedError(
                    "implicit RMHMC with jitter draws a data-dependent number of uniforms per iteration; its "
                    "reference stream cannot be pre-drawn -- use rng='philox'")
            J = (8 * L + 3) if jitter is not None else 0
            z = torch.empty((num_samples, D), dtype=torch.float32, device=q0.device)
            logu = torch.empty(num_samples, dtype=torch.float32)
            uni = torch.zeros((num_samples, max(J, 1), D), dtype=torch.float32)
            for n in range(num_samples):                     # fisher's rand(D) (:115) is always CPU, gibbs first (:184)
                if J:
                    uni[n, 0] = torch.rand(D)
                z[n] = torch.randn(D, dtype=torch.float32, device=q0.device)
                for j in range(1, J):
                    uni[n, j] = torch.rand(D)
                logu[n] = torch.log(torch.rand(1))[0]
            normals, log_uniforms = z.unsqueeze(1), logu.unsqueeze(1)
            uniforms = uni.unsqueeze(1) if J else None
        elif rng == 'injected':
            if normals is None or log_uniforms is None:
                raise RuntimeError("rng='injected' needs normals and log_uniforms")
            uniforms = injected_uniforms
        elif rng == 'philox':
            normals = log_uniforms = None
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)))
        else:
            raise ValueError('unknown rng mode %r' % (rng,))
        return engine.rmhmc_run(log_prob_func, q0, num_samples, L, step_size, burn=burn, jitter=jitter,
                                softabs_const=softabs_const, explicit_binding_const=explicit_binding_const,
                                fixed_point_threshold=fixed_point_threshold,
                                fixed_point_max_iterations=fixed_point_max_iterations,
                                jitter_max_tries=jitter_max_tries, explicit=explicit,
                                softabs=(metric == Metric.SOFTABS), jacdiag=jacdiag, seed=seed or 0,
                                chain_offset=chain_offset,
                                normals=normals, log_uniforms=log_uniforms, uniforms=uniforms, record_ham=record_ham)
    raise NotImplementedError()                                                                     # :606, :844


# ----------------------------------------------------------------------------------------------------------
# Bayesian neural networks: define_model_log_prob / sample_model / sample_split_model / predict_model
# ----------------------------------------------------------------------------------------------------------
def _check_loss(model_loss):
    if callable(model_loss) or model_loss not in T.LOSS_ID:
        raise NotImplementedError(
            "hamiltorch_b200: model_loss must be one of %s -- a callable loss (samplers.py:1186-1188) cannot enter a "
            "CUDA kernel and there is no CPU fallback; got %r" % (sorted(T.LOSS_ID), model_loss))


def define_model_log_prob(model, model_loss, x, y, params_flattened_list, params_shape_list, tau_list, tau_out,
                          normalizing_const=1., predict=False, prior_scale=1.0, device='cpu'):
    """samplers.py:1093-1201.  Returns the log_prob_func of a dense-stack model as a NATIVE descriptor
    (``targets.MLPRegression``): callable like the reference's closure (same torch ops, same (O,)-shaped value, the
    ``(log_prob, output)`` pair when ``predict``) and understood by the kernels."""
    _check_loss(model_loss)
    desc = T.MLPTarget.from_model(model, x, y, tau_list, tau_out, prior_scale, model_loss)
    if list(params_flattened_list) != desc.sizes:
        raise RuntimeError('params_flattened_list does not match the model')
    desc.predict_mode = bool(predict)
    return desc


def define_split_model_log_prob(model, model_loss, train_loader, num_splits, params_flattened_list, params_shape_list,
                                tau_list, tau_out, normalizing_const=1., predict=False, device='cpu', verbose=True):
    """samplers.py:1203-1258: one descriptor per DataLoader batch (the first ``num_splits`` batches), the prior divided
    by ``num_splits`` (:1254)."""
    log_prob_list = []
    for batch_idx, (data, target) in enumerate(train_loader):
        if batch_idx > num_splits - 1:
            break
        log_prob_list.append(define_model_log_prob(model, model_loss, data.clone().to('cpu'), target.clone().to('cpu'),
                                                   params_flattened_list, params_shape_list, tau_list, tau_out,
                                                   normalizing_const=normalizing_const, prior_scale=num_splits,
                                                   predict=predict, device=device))
    if verbose:
        print('Number of splits: ', len(log_prob_list), ' , each of batch size ', train_loader.batch_size, '\n')
    return log_prob_list


def _model_lists(model, tau_list):
    params_shape_list, params_flattened_list = [], []
    build_tau = tau_list is None
    if build_tau:
        tau_list = []
    for weights in model.parameters():                                                   # :1351-1355
        params_shape_list.append(weights.shape)
        params_flattened_list.append(weights.nelement())
        if build_tau:
            tau_list.append(torch.tensor(1.))
    return params_shape_list, params_flattened_list, tau_list


def sample_model(model, x, y, params_init, model_loss='multi_class_linear_output', num_samples=10,
                 num_steps_per_sample=10, step_size=0.1, burn=0, inv_mass=None, jitter=None, normalizing_const=1.,
                 softabs_const=None, explicit_binding_const=100, fixed_point_threshold=1e-5,
                 fixed_point_max_iterations=1000, jitter_max_tries=10, sampler=Sampler.HMC,
                 integrator=Integrator.IMPLICIT, metric=Metric.HESSIAN, debug=False, tau_out=1., tau_list=None,
                 store_on_GPU=True, desired_accept_rate=0.8, verbose=True, **engine_kwargs):
    """samplers.py:1261-1362: build the model's log_prob_func and delegate to ``sample``."""
    shapes, flat, tau_list = _model_lists(model, tau_list)
    log_prob_func = define_model_log_prob(model, model_loss, x, y, flat, shapes, tau_list, tau_out,
                                          normalizing_const=normalizing_const, device=params_init.device)
    return sample(log_prob_func, params_init, num_samples=num_samples, num_steps_per_sample=num_steps_per_sample,
                  step_size=step_size, burn=burn, jitter=jitter, inv_mass=inv_mass, normalizing_const=normalizing_const,
                  softabs_const=softabs_const, explicit_binding_const=explicit_binding_const,
                  fixed_point_threshold=fixed_point_threshold, fixed_point_max_iterations=fixed_point_max_iterations,
                  jitter_max_tries=jitter_max_tries, sampler=sampler, integrator=integrator, metric=metric, debug=debug,
                  desired_accept_rate=desired_accept_rate, store_on_GPU=store_on_GPU, verbose=verbose, **engine_kwargs)


def sample_split_model(model, train_loader, params_init, num_splits, model_loss='multi_class_linear_output',
                       num_samples=10, num_steps_per_sample=10, step_size=0.1, burn=0, inv_mass=None, jitter=None,
                       normalizing_const=1., softabs_const=None, explicit_binding_const=100,
                       fixed_point_threshold=1e-5, fixed_point_max_iterations=1000, jitter_max_tries=10,
                       sampler=Sampler.HMC, integrator=Integrator.SPLITTING, metric=Metric.HESSIAN, debug=False,
                       tau_out=1., tau_list=None, store_on_GPU=True, desired_accept_rate=0.8, verbose=True,
                       **engine_kwargs):
    """samplers.py:1364-1466: the first ``num_splits`` batches of ``train_loader`` become the data splits."""
    shapes, flat, tau_list = _model_lists(model, tau_list)
    log_prob_func = define_split_model_log_prob(model, model_loss, train_loader, num_splits, flat, shapes, tau_list,
                                                tau_out, normalizing_const=1., predict=False,
                                                device=params_init.device, verbose=verbose)
    return sample(log_prob_func, params_init, num_samples=num_samples, num_steps_per_sample=num_steps_per_sample,
                  step_size=step_size, burn=burn, jitter=jitter, inv_mass=inv_mass, normalizing_const=normalizing_const,
                  softabs_const=softabs_const, explicit_binding_const=explicit_binding_const,
                  fixed_point_threshold=fixed_point_threshold, fixed_point_max_iterations=fixed_point_max_iterations,
                  jitter_max_tries=jitter_max_tries, sampler=sampler, integrator=integrator, metric=metric, debug=debug,
                  desired_accept_rate=desired_accept_rate, store_on_GPU=store_on_GPU, verbose=verbose, **engine_kwargs)


def predict_model(model, samples, x=None, y=None, test_loader=None, model_loss='multi_class_linear_output', tau_out=1.,
                  tau_list=None, verbose=False):
    """samplers.py:1468-1562: network outputs of every sample as a (S, N, O) tensor + the list of S log-probs, from ONE
    launch (one CTA per sample).  With a DataLoader the batches are the splits and, like the reference (:1527), the
    prior is divided by the number of batches."""
    with torch.no_grad():
        shapes, flat, tau_list = _model_lists(model, tau_list)
        if test_loader.__class__ is torch.utils.data.dataloader.DataLoader:
            if len(test_loader.dataset) % test_loader.batch_size == 0.0:
                num_batches = len(test_loader.dataset) / test_loader.batch_size
            else:
                num_batches = int(round(len(test_loader.dataset) / test_loader.batch_size) + 1)
            target = define_split_model_log_prob(model, model_loss, test_loader, num_batches, flat, shapes, tau_list,
                                                 tau_out, normalizing_const=1., predict=True,
                                                 device=samples[0].device, verbose=verbose)
        elif x is not None and y is not None:
            if x.device != samples[0].device:
                raise RuntimeError('x on device: {} and samples on device: {}'.format(x.device, samples[0].device))
            target = define_model_log_prob(model, model_loss, x, y, flat, shapes, tau_list, tau_out, predict=True,
                                           device=samples[0].device)
        else:
            raise RuntimeError('Val data not defined (i.e. arguments x, y, val_loader are all not defined)')
        dev = samples[0].device
        stacked = torch.stack(list(samples))
        cap = N.MLP_MAX_SPLITS
        if isinstance(target, list) and len(target) > cap:
            # a loader with more batches than one launch takes splits (e.g. a 10k test set at batch_size 100): launches of
            # `cap` batches each, predictions concatenated in batch order, log-probs added (the reference's loop :1532-:1537)
            preds, lp = [], None
            for i in range(0, len(target), cap):
                pr, l = engine.mlp_predict(target[i:i + cap], stacked)
                preds.append(pr)
                lp = l if lp is None else lp + l
            pred = torch.cat(preds, dim=1)
        else:
            pred, lp = engine.mlp_predict(target, stacked)
        pred, lp = pred.to(dev), lp.to(dev)
    shape = (1,) if model_loss == 'regression' else ()          # the reference's closure: (O,) vs 0-d
    return pred, [l.reshape(shape) for l in lp.unbind(0)]
