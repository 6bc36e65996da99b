"""Host-side driver of the sm_100a kernels: turns descriptors + torch tensors into C-ABI calls (include/hmcx.h).

Everything here is plumbing -- device buffers, padding to the (C, ld) layout, the step-size adaptation table,
the random-stream modes.  The arithmetic of the hot path lives in csrc/*.cu.
"""
import ctypes as C
import os
import math

import torch

from . import _native as N
from . import targets as T


# ----------------------------------------------------------------------------------------------------------
# descriptors -> native structs (tensors are kept alive on the wrapper object)
# ----------------------------------------------------------------------------------------------------------
def _combine_mlp_splits(splits):
    """A list of MLPRegression descriptors (what define_split_model_log_prob returns: one per data batch, sharing the
    network and the prior) -> (first descriptor, concatenated x, y, split_begin)."""
    first = splits[0]
    for d in splits:
        if not isinstance(d, T.MLPRegression):
            raise TypeError('a split log_prob_func list must hold MLPRegression descriptors')
        if d.widths != first.widths or d.acts != first.acts or d.tau_out != first.tau_out or \
                d.model_loss != first.model_loss or \
                float(d.prior_scale) != float(first.prior_scale) or \
                any(float(a) != float(b) for a, b in zip(d.tau_list, first.tau_list)):
            raise RuntimeError('all splits must share the network, tau_list, tau_out and prior_scale')
        if d.x is None:
            raise RuntimeError('split descriptors need data')
    x = torch.cat([d.x for d in splits])
    y = torch.cat([d.y.reshape(d.x.shape[0], first.y_cols) for d in splits])
    begin = [0]
    for d in splits:
        begin.append(begin[-1] + d.x.shape[0])
    return first, x, y, begin


class NativeTarget:
    def __init__(self, target, device):
        self.device = torch.device(device)
        self._keep = {}
        if isinstance(target, list) or isinstance(target, T.MLPRegression):
            self._init_mlp(target)
            return
        if not T.is_target(target):
            raise TypeError('not a hamiltorch_b200 target descriptor: %r' % (target,))
        self.target = target
        self.dim = target.dim
        self.num_splits = 1
        s = N.TargetStruct()
        s.kind, s.dim = target.kind, target.dim
        s.log_norm = float(getattr(target, 'log_norm', 0.0))
        for field, attr in (('mean', 'mean'), ('inv_var', 'inv_var'), ('prec', 'prec')):
            t = getattr(target, attr, None)
            if t is not None:
                t = t.detach().to(self.device, torch.float32).contiguous()
                self._keep[field] = t
                setattr(s, field, t.data_ptr())
        s.funnel_inv_var_v = float(getattr(target, 'inv_var_v', 0.0))
        self.struct = s

    def _init_mlp(self, target):
        if isinstance(target, list):
            first, x, y, begin = _combine_mlp_splits(target)
        else:
            first = target
            x = first.x
            y = None if x is None else first.y.reshape(x.shape[0], first.y_cols)
            begin = [0, 0 if x is None else x.shape[0]]
        if len(begin) - 1 > N.MLP_MAX_SPLITS:
            raise NotImplementedError('at most %d splits' % N.MLP_MAX_SPLITS)
        self.target = target
        self.dim = first.dim
        self.num_splits = len(begin) - 1
        self.mlp_desc = first
        m = N.MlpStruct()
        m.num_layers = first.num_layers
        for i, w in enumerate(first.widths):
            m.widths[i] = w
        for i, a in enumerate(first.acts):
            m.activation[i] = a
        m.loss = first.loss_id
        m.tau_out = first.tau_out
        m.prior_scale = float(first.prior_scale)
        for i in range(2 * first.num_layers):
            m.prior_two_var[i] = float(first.two_var[i])
            m.prior_log_scale[i] = float(first.log_scale[i])
            m.prior_grad_coef[i] = float(first.grad_coef[i])
        if x is not None:
            xd = x.detach().to(self.device, torch.float32).contiguous()
            yd = y.detach().to(self.device, torch.float32).contiguous()
            self._keep['x'], self._keep['y'] = xd, yd
            m.x, m.y = xd.data_ptr(), yd.data_ptr()
            m.num_rows = xd.shape[0]
        m.num_splits = self.num_splits
        m.cluster_size = int(getattr(first, 'cluster_size', 0))      # 0 = auto; 1/2/4 pins CTAs per chain
        m.tensor_cores = int(getattr(first, 'tensor_cores', 0))      # 0 = auto (tcgen05 when the shape fits); 1 = off
        for i, b in enumerate(begin):
            m.split_begin[i] = b
        self.mlp_struct = m
        s = N.TargetStruct()
        s.kind, s.dim = T.KIND_MLP, first.dim
        s.mlp = C.pointer(m)
        self.struct = s
        if x is not None and m.tensor_cores == 0 and self.device.type == 'cuda':
            # tensor-core form: x as ready-made tcgen05 operands (tf32 hi | lo, both GEMM layouts), built once per target
            lib = N.load_library()
            nbytes = int(lib.hmcx_mlp_packed_x_bytes(C.byref(s)))
            if nbytes:
                xp = torch.empty(nbytes // 4, dtype=torch.float32, device=self.device)
                with torch.cuda.device(self.device):
                    N.check(lib.hmcx_mlp_pack_x(C.byref(s), N.ptr(xp), N.stream_ptr(self.device)), 'hmcx_mlp_pack_x')
                self._keep['x_packed'] = xp
                m.x_packed = xp.data_ptr()

    def ref(self):
        return C.byref(self.struct)


# Descriptor targets (Gaussian / Funnel) carry small host tensors; uploading them is a pageable host-to-device copy = a host
# synchronisation per call.  The same descriptor object, unmodified, reuses its device operands (keyed like _MASS_CACHE; the
# entry holds the descriptor).  Bayesian-NN targets (lists / MLPRegression: data + packed operands) are rebuilt per call.
_TARGET_CACHE = []


def native_target(target, device):
    if isinstance(target, NativeTarget):
        return target
    if isinstance(target, list) or isinstance(target, T.MLPRegression) or not T.is_target(target):
        return NativeTarget(target, device)
    tensors = [t for t in (getattr(target, a, None) for a in ('mean', 'inv_var', 'prec')) if torch.is_tensor(t)]
    key = (id(target), tuple(id(t) for t in tensors), tuple(t._version for t in tensors), str(torch.device(device)),
           target.dim, float(getattr(target, 'log_norm', 0.0)), float(getattr(target, 'inv_var_v', 0.0)))
    for k, _, nt in _TARGET_CACHE:
        if k == key:
            return nt
    nt = NativeTarget(target, device)
    _TARGET_CACHE.append((key, target, nt))
    if len(_TARGET_CACHE) > 8:
        _TARGET_CACHE.pop(0)
    return nt


class NativeMass:
    """inv_mass as the reference accepts it (None | (D,) | (D,D)); the mass used by gibbs is inverted ONCE with
    the same torch ops as samplers.py:942-952 so that sqrt(mass) is bit-identical."""

    def __init__(self, inv_mass, dim, device):
        self.device = torch.device(device)
        s = N.MassStruct()
        self._keep = {}
        if inv_mass is None:
            s.kind = N.MASS_NONE
        elif isinstance(inv_mass, list):
            # block list (samplers.py:188-197, :287-292, :803-809, :944-947) == the block-diagonal 2-D inv_mass: every
            # block inverted and Cholesky-factorised on its own with the reference's torch ops, then laid out as ONE
            # (D, D) operand pair for the full-mass kernels (thread-per-chain for D <= 16, tcgen05 dense_lin above)
            blocks = [b.detach().to(torch.float32) for b in inv_mass]
            if any(b.dim() != 2 or b.shape[0] != b.shape[1] for b in blocks) or sum(b.shape[0] for b in blocks) != dim:
                raise RuntimeError('block-list inv_mass: square blocks whose sizes add up to %d' % dim)
            im = torch.block_diag(*blocks)
            tril = torch.block_diag(*[torch.linalg.cholesky(torch.inverse(b)) for b in blocks])
            s.kind = N.MASS_FULL
            self._keep['im'] = im.to(self.device).contiguous()
            self._keep['tril'] = tril.to(self.device).contiguous()
            s.inv_mass = self._keep['im'].data_ptr()
            s.mass_factor = self._keep['tril'].data_ptr()
        elif inv_mass.dim() == 1:
            if inv_mass.numel() != dim:
                raise RuntimeError('inv_mass must have %d entries' % dim)
            im = inv_mass.detach().to(torch.float32)
            sd = (1 / im) ** 0.5                       # mass = 1/inv_mass (:952); Normal(0, mass**0.5) (:201)
            s.kind = N.MASS_DIAG
            self._keep['im'] = im.to(self.device).contiguous()
            self._keep['sd'] = sd.to(self.device).contiguous()
            s.inv_mass = self._keep['im'].data_ptr()
            s.mass_factor = self._keep['sd'].data_ptr()
        elif inv_mass.dim() == 2:
            if tuple(inv_mass.shape) != (dim, dim):
                raise RuntimeError('inv_mass must be (%d, %d)' % (dim, dim))
            im = inv_mass.detach().to(torch.float32)
            mass = torch.inverse(im)                   # :950
            tril = torch.linalg.cholesky(mass)         # MultivariateNormal(0, mass).scale_tril (:199)
            s.kind = N.MASS_FULL
            self._keep['im'] = im.to(self.device).contiguous()
            self._keep['tril'] = tril.to(self.device).contiguous()
            s.inv_mass = self._keep['im'].data_ptr()
            s.mass_factor = self._keep['tril'].data_ptr()
        else:
            raise RuntimeError('inv_mass must be None, 1-D or 2-D')
        self.struct = s

    @property
    def kind(self):
        return self.struct.kind

    def ref(self):
        return C.byref(self.struct)


# A full (2-D) or block-list inv_mass costs a host inversion + Cholesky factorisation (samplers.py:942-952 does it once per
# sample() call).  Repeated calls with the SAME tensor object, unmodified (torch's version counter), reuse the device
# operands: the entry holds a reference to the caller's tensor(s), so a recycled address can never alias it.
_MASS_CACHE = []


def native_mass(inv_mass, dim, device):
    if isinstance(inv_mass, NativeMass):
        return inv_mass
    heavy = isinstance(inv_mass, list) or (torch.is_tensor(inv_mass) and inv_mass.dim() == 2)
    if not heavy:
        return NativeMass(inv_mass, dim, device)
    parts = inv_mass if isinstance(inv_mass, list) else [inv_mass]
    key = (tuple(id(t) for t in parts), tuple(t._version for t in parts), dim, str(torch.device(device)))
    for k, refs, nm in _MASS_CACHE:
        if k == key:
            return nm
    nm = NativeMass(inv_mass, dim, device)
    _MASS_CACHE.append((key, list(parts), nm))
    if len(_MASS_CACHE) > 4:
        _MASS_CACHE.pop(0)
    return nm


_SIDE_STREAMS = {}


def _side_stream(device):
    key = str(torch.device(device))
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def nuts_table(burn):
    """The five Python-double constants of samplers.py:659-671 for t = 1..burn+1 (gamma=.05, t0=10, kappa=.75)."""
    rows = []
    for n in range(burn + 1):
        t = n + 1
        w = (1 / (t + 10))
        rows.append([1 - w, w, (t ** 0.5) / 0.05, t ** -0.75, 1 - t ** -0.75])
    return torch.tensor(rows, dtype=torch.float64)


_NUTS_TABLES = {}


def nuts_table_device(burn, device):
    """The constants of nuts_table() on the device, built and uploaded once per (burn, device): the upload is a pageable
    host-to-device copy, i.e. a host synchronisation with everything queued on the stream -- per call it kept a sampler that
    is called in a loop from ever running ahead of the GPU."""
    key = (int(burn), str(torch.device(device)))
    t = _NUTS_TABLES.get(key)
    if t is None:
        if len(_NUTS_TABLES) > 16:
            _NUTS_TABLES.clear()
        t = _NUTS_TABLES[key] = nuts_table(burn).to(device)
    return t


def nuts_mu(step_size_init):
    """samplers.py:664 -- fp32 log of fp32(10*eps0), returned as a Python float."""
    return float(torch.log(10 * torch.FloatTensor([step_size_init])))


# ----------------------------------------------------------------------------------------------------------
# kernel entry points on torch tensors
# ----------------------------------------------------------------------------------------------------------
def _as_rows(x, ld, device):
    x = x.detach().to(device=device, dtype=torch.float32)
    if x.dim() == 1:
        x = x.unsqueeze(0)
    return N.pad_rows(x.contiguous(), ld)


def _eps_vector(step_size, C, device):
    if torch.is_tensor(step_size):
        e = step_size.detach().to(device=device, dtype=torch.float32).reshape(-1)
        if e.numel() == 1:
            e = e.expand(C)
        return e.contiguous().clone()
    return torch.full((C,), float(step_size), dtype=torch.float32, device=device)


def leapfrog(target, q, p, steps, step_size, inv_mass=None, return_trajectory=False, device=None):
    """Batched samplers.leapfrog (plain HMC branch).  q, p: (C, D) or (D,).  Returns (q_L, p_L) as (C, D), or
    the (L, C, D) trajectories when ``return_trajectory``."""
    N.require_cuda()
    lib = N.load_library()
    device = torch.device(device if device is not None else (q.device if q.is_cuda else 'cuda'))
    nt = native_target(target, device)
    D, ld = nt.dim, N.padded_ld(nt.dim)
    nm = native_mass(inv_mass, D, device)
    qd, pd = _as_rows(q, ld, device), _as_rows(p, ld, device)
    Cn = qd.shape[0]
    eps = _eps_vector(step_size, Cn, device)
    q_out, p_out = torch.empty_like(qd), torch.empty_like(pd)
    q_traj = p_traj = None
    if return_trajectory:
        q_traj = torch.empty((steps, Cn, ld), dtype=torch.float32, device=device)
        p_traj = torch.empty_like(q_traj)
    with torch.cuda.device(device):
        rc = lib.hmcx_leapfrog(nt.ref(), nm.ref(), N.ptr(qd), N.ptr(pd), N.ptr(eps), Cn, ld, int(steps),
                               N.ptr(q_out), N.ptr(p_out), N.ptr(q_traj), N.ptr(p_traj), N.stream_ptr(device))
    N.check(rc, 'hmcx_leapfrog')
    if return_trajectory:
        return q_traj[..., :D], p_traj[..., :D]
    return q_out[:, :D], p_out[:, :D]


def split_leapfrog(targets, q, p, steps, step_size, scheme, inv_mass=None, perms=None, seed=0, device=None):
    """Batched samplers.leapfrog with Integrator.SPLITTING / SPLITTING_RAND / SPLITTING_KMID (:494-603) on the list of
    data-split closures ``targets``.  q, p: (C, D) or (D,).  Returns the (L, C, D) trajectories of params and momentum
    (the state after every step).  SPLITTING_RAND: ``perms`` (C, M) injects the call's randperm(M) (:550), else Philox."""
    N.require_cuda()
    lib = N.load_library()
    device = torch.device(device if device is not None else (q.device if q.is_cuda else 'cuda'))
    nt = native_target(targets, device)
    D, ld = nt.dim, N.padded_ld(nt.dim)
    if isinstance(inv_mass, list) or (torch.is_tensor(inv_mass) and inv_mass.dim() != 1):
        raise NotImplementedError('split leapfrog: inv_mass None or 1-D')
    nm = native_mass(inv_mass, D, device)
    qd, pd = _as_rows(q, ld, device), _as_rows(p, ld, device)
    Cn, L = qd.shape[0], int(steps)
    eps = _eps_vector(step_size, Cn, device)
    rng = N.RngStruct()
    keep = []
    if perms is not None:
        pm = perms.detach().to(device=device, dtype=torch.int32).reshape(Cn, nt.num_splits).contiguous()
        rng.mode, rng.perms = N.RNG_INJECTED, pm.data_ptr()
        keep.append(pm)
    else:
        rng.mode, rng.seed = N.RNG_PHILOX, int(seed)
    q_traj = torch.zeros((L, Cn, ld), dtype=torch.float32, device=device)
    p_traj = torch.zeros_like(q_traj)
    with torch.cuda.device(device):
        rc = lib.hmcx_split_leapfrog(nt.ref(), nm.ref(), C.byref(rng), int(scheme), float(step_size), N.ptr(qd), N.ptr(pd),
                                     N.ptr(eps), Cn, ld, L, N.ptr(q_traj), N.ptr(p_traj), N.stream_ptr(device))
    N.check(rc, 'hmcx_split_leapfrog')
    torch.cuda.current_stream(device).synchronize()
    return q_traj[..., :D], p_traj[..., :D]


def hamiltonian(target, q, p, inv_mass=None, device=None):
    """Batched samplers.hamiltonian (sampler=HMC).  Returns (H (C,), nonfinite_flags (C,) uint8)."""
    N.require_cuda()
    lib = N.load_library()
    device = torch.device(device if device is not None else (q.device if q.is_cuda else 'cuda'))
    nt = native_target(target, device)
    D, ld = nt.dim, N.padded_ld(nt.dim)
    nm = native_mass(inv_mass, D, device)
    qd, pd = _as_rows(q, ld, device), _as_rows(p, ld, device)
    Cn = qd.shape[0]
    H = torch.empty(Cn, dtype=torch.float32, device=device)
    flags = torch.empty(Cn, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        rc = lib.hmcx_hamiltonian(nt.ref(), nm.ref(), N.ptr(qd), N.ptr(pd), Cn, ld, N.ptr(H), N.ptr(flags),
                                  N.stream_ptr(device))
    N.check(rc, 'hmcx_hamiltonian')
    return H, flags


def gibbs(dim, num_chains, seed, iteration=0, inv_mass=None, chain_offset=0, device='cuda'):
    """Batched samplers.gibbs (sampler=HMC) from the in-kernel Philox stream."""
    N.require_cuda()
    lib = N.load_library()
    device = torch.device(device)
    ld = N.padded_ld(dim)
    nm = native_mass(inv_mass, dim, device)
    rng = N.RngStruct()
    rng.mode, rng.seed, rng.chain_offset = N.RNG_PHILOX, int(seed), int(chain_offset)
    p = torch.empty((num_chains, ld), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        rc = lib.hmcx_gibbs(nm.ref(), C.byref(rng), dim, num_chains, ld, int(iteration), N.ptr(p),
                            N.stream_ptr(device))
    N.check(rc, 'hmcx_gibbs')
    return p[:, :dim]


class HMCResult:
    """Device-resident result of a batched run."""

    def __init__(self, samples, accepted, diverged, ham, step_size, num_rejected, dim, num_samples):
        self.samples_padded = samples            # (C, S-burn, ld)
        self.dim = dim
        self.accepted = accepted                 # (C, S) uint8
        self.diverged = diverged                 # (C, S) uint8
        self.ham = ham                           # (C, S, 2) or None
        self.step_size = step_size               # (C,) final per-chain eps
        self.num_rejected = num_rejected         # (C,) int32
        self.num_samples = num_samples

    @property
    def samples(self):
        if self.samples_padded is None:
            raise RuntimeError('this run kept no samples (keep_samples=False): use moment_sum / moment_sumsq')
        return self.samples_padded[..., :self.dim]

    @property
    def accept_rate(self):
        return 1.0 - self.num_rejected.double() / self.num_samples


def hmc_run(target, params_init, num_samples, num_steps_per_sample, step_size, burn=0, inv_mass=None,
            nuts=False, desired_accept_rate=0.8, seed=0, chain_offset=0, normals=None, log_uniforms=None,
            record_ham=False, out=None, device=None, tuning=0, eps_schedule=None, record_eps=False, scheme=None,
            perms=None, thin=1, moments=False, keep_samples=True, host_samples=False, host_windows=0):
    """The reference's sample() loop for sampler in {HMC, HMC_NUTS} as one persistent kernel over C chains.

    params_init (C, D) | (D,).  Randomness: in-kernel Philox keyed by (seed, chain_offset+c, iteration), or -- when
    ``normals`` (S, C, D) and ``log_uniforms`` (S, C) are given -- the injected stream (parity mode).
    ``out``: optional pre-allocated (C, S-burn, ld) fp32 tensor for the samples: a device tensor, or a PINNED host
    tensor (the kernel then streams the retained rows straight into it, see ``host_samples``).
    Bayesian-NN targets (an MLPRegression or the list of split descriptors): ``scheme`` selects the integrator
    (N.SCHEME_PLAIN / SPLIT_SYM / SPLIT_RAND / SPLIT_KMID); ``perms`` (S, C, M) injects SPLITTING_RAND's randperm.
    NUTS only: ``eps_schedule`` (S, C) forces the step size of every iteration (parity tests replay the reference's
    schedule); ``record_eps`` returns the kernel's own adapted step sizes in ``result.eps_trace`` (C, S).
    Sample sink (include/hmcx.h hmcx_sink_t; element-wise targets): ``thin`` keeps every thin-th post-burn state,
    ``moments`` accumulates per-chain running sum / sum of squares over every post-burn iteration in the kernel's
    registers with compensated (Neumaier) summation (``result.moment_sum``, ``result.moment_sumsq``: fp64 tensors
    relative error ~ n*eps^2 instead of the naive n*eps; ``result.moment_count``), ``keep_samples=False`` stores no
    samples at all, ``host_samples=True`` makes the kernel stream the retained rows straight into pinned host memory
    (the reference's ``store_on_GPU=False``, samplers.py:1008-1012) -- ``result.samples`` is then a CPU tensor, valid
    after a stream synchronisation.
    """
    N.require_cuda()
    lib = N.load_library()
    if device is None:
        device = params_init.device if params_init.is_cuda else torch.device('cuda', torch.cuda.current_device())
    device = torch.device(device)
    nt = native_target(target, device)
    D, ld = nt.dim, N.padded_ld(nt.dim)
    nm = native_mass(inv_mass, D, device)
    S, L, burn = int(num_samples), int(num_steps_per_sample), int(burn)
    q_init = _as_rows(params_init, ld, device)
    if q_init.shape[1] != ld or params_init.shape[-1] != D:
        raise RuntimeError('params_init last dimension must be %d' % D)
    Cn = q_init.shape[0]
    q_cur = q_init.clone()
    eps = _eps_vector(step_size, Cn, device)
    thin = int(thin)
    if thin < 1:
        raise RuntimeError('thin must be >= 1')
    host_out = None
    if out is not None and not out.is_cuda:
        if not out.is_pinned():
            raise RuntimeError('a host `out` buffer must be pinned (page-locked) memory')
        if (int(host_windows) >= 2 and thin == 1 and not moments and keep_samples and scheme is None and ld <= 4096 and
                nt.struct.kind in (T.GaussianIso.kind, T.GaussianDiag.kind) and nm.kind != N.MASS_FULL and normals is None):
            # windowed delivery: the run is cut into `host_windows` windows of iterations; each window's sample slots
            # leave for the pinned block through the COPY ENGINE on a second stream while the next window computes
            # (hmcx_copy_rows_async).  Costs a device staging block of the samples' size; delivers at the DMA rate
            # (57 GB/s on B200/PCIe 5) where SM-issued stores to host memory reach ~52.5.
            host_out, out = out, None
        else:
            host_samples = True                    # caller-provided pinned sample block: the kernel streams into it
    use_sink = thin > 1 or moments or not keep_samples or host_samples
    if use_sink and scheme is not None:
        raise NotImplementedError('the sample sink is implemented for element-wise targets')
    keep = 1 + (S - burn - 1) // thin
    if not keep_samples:
        samples = None
    elif host_samples and out is None:
        # pinned host memory is device-addressable under unified virtual addressing: the kernel's st.global.cs rows go
        # over PCIe while the chains keep running (no device-side sample buffer, no separate D2H copy)
        samples = torch.empty((Cn, keep, ld), dtype=torch.float32, pin_memory=True)
    elif out is None:
        samples = torch.empty((Cn, keep, ld), dtype=torch.float32, device=device)
    else:
        samples = out
        if tuple(samples.shape) != (Cn, keep, ld) or samples.dtype != torch.float32 or not samples.is_contiguous():
            raise RuntimeError('out must be a contiguous fp32 (C, S-burn, ld) tensor')
    accepted = torch.empty((Cn, S), dtype=torch.uint8, device=device)
    diverged = torch.empty((Cn, S), dtype=torch.uint8, device=device)
    ham = torch.empty((Cn, S, 2), dtype=torch.float32, device=device) if record_ham else None
    num_rejected = torch.zeros(Cn, dtype=torch.int32, device=device)

    rng = N.RngStruct()
    keep_alive = []
    if normals is not None:
        z = normals.detach().to(device=device, dtype=torch.float32)
        if z.dim() == 2:
            z = z.unsqueeze(1)
        if tuple(z.shape[:2]) != (S, Cn) or z.shape[2] != D:
            raise RuntimeError('normals must be (S, C, D)')
        z = N.pad_rows(z.contiguous(), ld)
        lu = log_uniforms.detach().to(device=device, dtype=torch.float32).reshape(S, Cn).contiguous()
        rng.mode = N.RNG_INJECTED
        rng.normals, rng.log_uniforms = z.data_ptr(), lu.data_ptr()
        keep_alive += [z, lu]
        if perms is not None:
            if perms.dim() != 3 or tuple(perms.shape[:2]) != (S, Cn) or perms.shape[2] != nt.num_splits:
                raise RuntimeError('perms must be (S, C, M) = (%d, %d, %d), got %s'
                                   % (S, Cn, nt.num_splits, tuple(perms.shape)))
            pm = perms.detach().to(device=device, dtype=torch.int32).contiguous()
            rng.perms = pm.data_ptr()
            keep_alive.append(pm)
    else:
        rng.mode, rng.seed, rng.chain_offset = N.RNG_PHILOX, int(seed), int(chain_offset)

    nuts_s = N.NutsStruct()
    eps_trace = None
    if not torch.is_tensor(step_size):
        nuts_s.step_size_init = float(step_size)           # the double the reference divides in its split drifts
    if nuts:
        table = nuts_table_device(burn, device)
        h_bar = torch.zeros(Cn, dtype=torch.float64, device=device)
        eps_bar = torch.ones(Cn, dtype=torch.float64, device=device)
        nuts_s.enabled = 1
        nuts_s.desired_accept_rate = float(desired_accept_rate)
        nuts_s.mu = nuts_mu(step_size if not torch.is_tensor(step_size) else float(step_size.reshape(-1)[0]))
        nuts_s.table, nuts_s.h_bar, nuts_s.eps_bar = table.data_ptr(), h_bar.data_ptr(), eps_bar.data_ptr()
        keep_alive += [table, h_bar, eps_bar]
        if eps_schedule is not None:
            sched = eps_schedule.detach().to(device=device, dtype=torch.float32).reshape(S, Cn).contiguous()
            nuts_s.eps_schedule = sched.data_ptr()
            keep_alive.append(sched)
        if record_eps:
            eps_trace = torch.zeros((Cn, S), dtype=torch.float32, device=device)
            nuts_s.eps_trace = eps_trace.data_ptr()

    msum = msq = msum_lo = msq_lo = None
    with torch.cuda.device(device):
        if scheme is None:
            ws_bytes = lib.hmcx_hmc_workspace_bytes(nt.ref(), nm.ref(), Cn, ld)
            ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=device) if ws_bytes else None
            if use_sink:
                sink = N.SinkStruct()
                sink.thin = thin
                if moments:
                    msum, msq, msum_lo, msq_lo = (torch.zeros((Cn, ld), dtype=torch.float32, device=device)
                                                  for _ in range(4))
                    sink.sum, sink.sumsq = msum.data_ptr(), msq.data_ptr()
                    sink.sum_lo, sink.sumsq_lo = msum_lo.data_ptr(), msq_lo.data_ptr()
                rc = lib.hmcx_hmc_run_sink(nt.ref(), nm.ref(), C.byref(rng), C.byref(nuts_s), N.ptr(q_init),
                                           N.ptr(q_cur), N.ptr(eps), Cn, ld, L, S, burn, 0, S, N.ptr(samples),
                                           N.ptr(accepted), N.ptr(diverged), N.ptr(ham), N.ptr(num_rejected),
                                           int(tuning), N.ptr(ws), C.byref(sink), N.stream_ptr(device))
                N.check(rc, 'hmcx_hmc_run_sink')
            elif host_out is not None:
                if tuple(host_out.shape) != (Cn, keep, ld) or host_out.dtype != torch.float32 or not host_out.is_contiguous():
                    raise RuntimeError('out must be a contiguous fp32 (C, S-burn, ld) tensor')
                W = min(int(host_windows), S)
                main = torch.cuda.current_stream(device)
                side = _side_stream(device)
                pitch, slot_bytes = keep * ld * 4, ld * 4
                for w in range(W):
                    a, b = (S * w) // W, (S * (w + 1)) // W
                    rc = lib.hmcx_hmc_run(nt.ref(), nm.ref(), C.byref(rng), C.byref(nuts_s), N.ptr(q_init), N.ptr(q_cur),
                                          N.ptr(eps), Cn, ld, L, S, burn, a, b, N.ptr(samples), N.ptr(accepted),
                                          N.ptr(diverged), N.ptr(ham), N.ptr(num_rejected), int(tuning), N.ptr(ws),
                                          C.c_void_p(main.cuda_stream))
                    N.check(rc, 'hmcx_hmc_run')
                    lo = 0 if a == 0 else max(a - burn, 1)               # slots this window wrote: 0 = params_init (:959),
                    hi = max(b - burn, 1 if a == 0 else lo)              # n - burn for every iteration n > burn
                    if hi > lo:
                        ev = torch.cuda.Event()
                        ev.record(main)
                        side.wait_event(ev)
                        rc = lib.hmcx_copy_rows_async(C.c_void_p(host_out.data_ptr() + lo * slot_bytes), pitch,
                                                      C.c_void_p(samples.data_ptr() + lo * slot_bytes), pitch,
                                                      (hi - lo) * slot_bytes, Cn, C.c_void_p(side.cuda_stream))
                        N.check(rc, 'hmcx_copy_rows_async')
                done = torch.cuda.Event()
                done.record(side)
                main.wait_event(done)                       # the caller's stream sees the delivered block
                keep_alive.append(samples)
                samples = host_out
            else:
                rc = lib.hmcx_hmc_run(nt.ref(), nm.ref(), C.byref(rng), C.byref(nuts_s), N.ptr(q_init), N.ptr(q_cur),
                                      N.ptr(eps), Cn, ld, L, S, burn, 0, S, N.ptr(samples), N.ptr(accepted),
                                      N.ptr(diverged), N.ptr(ham), N.ptr(num_rejected), int(tuning), N.ptr(ws),
                                      N.stream_ptr(device))
                N.check(rc, 'hmcx_hmc_run')
        else:
            rc = lib.hmcx_split_run(nt.ref(), nm.ref(), C.byref(rng), C.byref(nuts_s), int(scheme), N.ptr(q_init),
                                    N.ptr(q_cur), N.ptr(eps), Cn, ld, L, S, burn, 0, S, N.ptr(samples),
                                    N.ptr(accepted), N.ptr(diverged), N.ptr(ham), N.ptr(num_rejected),
                                    N.stream_ptr(device))
            N.check(rc, 'hmcx_split_run')
    res = HMCResult(samples, accepted, diverged, ham, eps, num_rejected, D, S)
    res.eps_trace = eps_trace
    res.thin = thin
    # compensated running sums: hi + lo combined in fp64 (relative error ~2^-46 whatever the run length)
    res.moment_sum = None if msum is None else (msum.double() + msum_lo.double())[:, :D]
    res.moment_sumsq = None if msq is None else (msq.double() + msq_lo.double())[:, :D]
    res.moment_count = S - burn - 1
    res.final_state = q_cur[:, :D]
    if nuts:
        res.eps_bar, res.h_bar = eps_bar, h_bar
    res._keep_alive = keep_alive          # buffers the asynchronous kernel still reads
    return res


def grad_log_prob(target, q, split=-1, want_grad=True, want_log_prob=True, device=None):
    """Batched collect_gradients (samplers.py:33-66): (grad (C, D), log_prob (C,)) of C parameter vectors for a
    Bayesian-NN target; ``split`` selects one data split (-1: the whole potential)."""
    N.require_cuda()
    lib = N.load_library()
    device = torch.device(device if device is not None else (q.device if q.is_cuda else 'cuda'))
    nt = native_target(target, device)
    D, ld = nt.dim, N.padded_ld(nt.dim)
    qd = _as_rows(q, ld, device)
    Cn = qd.shape[0]
    g = torch.empty_like(qd) if want_grad else None
    lp = torch.empty(Cn, dtype=torch.float32, device=device) if want_log_prob else None
    with torch.cuda.device(device):
        rc = lib.hmcx_grad_log_prob(nt.ref(), N.ptr(qd), Cn, ld, int(split), N.ptr(g), N.ptr(lp), N.stream_ptr(device))
    N.check(rc, 'hmcx_grad_log_prob')
    return (g[:, :D] if want_grad else None), lp


def mlp_predict(target, samples, device=None):
    """Batched predict_model (samplers.py:1468-1562): samples (S, D) -> (pred (S, N, O), log_prob (S,))."""
    N.require_cuda()
    lib = N.load_library()
    device = torch.device(device if device is not None else (samples.device if samples.is_cuda else 'cuda'))
    nt = native_target(target, device)
    D, ld = nt.dim, N.padded_ld(nt.dim)
    sd = _as_rows(samples, ld, device)
    S = sd.shape[0]
    m = nt.mlp_struct
    pred = torch.empty((S, m.num_rows, m.widths[m.num_layers]), dtype=torch.float32, device=device)   # logits / log-probs
    lp = torch.empty(S, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        rc = lib.hmcx_mlp_predict(nt.ref(), N.ptr(sd), S, ld, N.ptr(pred), N.ptr(lp), N.stream_ptr(device))
    N.check(rc, 'hmcx_mlp_predict')
    return pred, lp


def const_metric(target, softabs, softabs_const):
    """The metric of a Gaussian target without jitter is one matrix: evaluate it ONCE on the host with the reference's
    own torch ops (fisher, samplers.py:108-121: autograd Hessian of the descriptor, eigh + coth map for SOFTABS) and
    derive what the kernels consume: G^-1 (the metric solve of cholesky_inverse :146-148 as a matrix), chol(G) (gibbs
    :183-184 through MultivariateNormal's scale_tril) and log det G (:726 / :728)."""
    D = target.dim
    # -Hessian of the quadratic log-density; what torch.autograd.functional.hessian(target, q) (:108) returns at ANY q
    # (tests/test_targets.py::test_const_metric_matches_the_oracle_fisher) without its D backward passes
    if isinstance(target, T.GaussianFull):
        fish = target.prec.detach().clone()
    elif isinstance(target, T.GaussianDiag):
        fish = torch.diag(target.inv_var.detach().to(torch.float32))
    elif isinstance(target, T.GaussianIso):
        fish = torch.eye(D, dtype=torch.float32)
    else:
        fish = -torch.autograd.functional.hessian(target, torch.zeros(D), create_graph=False)
    if softabs:
        lam, vec = torch.linalg.eigh(fish, UPLO='L')
        abs_lam = (1. / torch.tanh(softabs_const * lam)) * lam
        fish = torch.matmul(vec, torch.matmul(abs_lam.diag(), vec.t()))
        log_det = float(abs_lam.log().sum())
    else:
        log_det = float(torch.slogdet(fish)[1])
    lower = torch.linalg.cholesky(fish)
    ginv = torch.cholesky_inverse(lower.double()).float()        # (L L^T)^-1 from the reference's fp32 factor
    return ginv.contiguous(), lower.contiguous(), log_det


# The constant metric costs an eigh / Cholesky / inverse on the host (the reference pays them at EVERY fisher() call).  Repeated
# runs on the SAME target object, unmodified, reuse the device operands -- keyed like _MASS_CACHE; the entry holds the target.
_METRIC_CACHE = []


def const_metric_device(target, softabs, softabs_const, device):
    tensors = [t for t in (getattr(target, 'prec', None), getattr(target, 'inv_var', None)) if torch.is_tensor(t)]
    key = (id(target), tuple(id(t) for t in tensors), tuple(t._version for t in tensors), bool(softabs),
           float(softabs_const) if softabs else None, str(torch.device(device)))
    for k, _, val in _METRIC_CACHE:
        if k == key:
            return val
    ginv, lower, log_det = const_metric(target, softabs, softabs_const)
    val = (ginv.to(device), lower.to(device), log_det)
    _METRIC_CACHE.append((key, target, val))
    if len(_METRIC_CACHE) > 4:
        _METRIC_CACHE.pop(0)
    return val


def _rmhmc_is_dense(target, jitter, jacdiag=False):
    """Gaussian targets without jitter have a constant metric: the tensor-core path (GaussianFull at any D; GaussianIso /
    GaussianDiag above D = 16, below they stay on the thread-per-chain kernel).  Everything else -- Funnel, jitter,
    Metric.JACOBIAN_DIAG (depends on the gradient, i.e. on the position) -- assembles and factorises its metric inside
    the kernel: hmcx_rmhmc_run (D <= 16 one thread per chain, D <= 64 one CTA per chain)."""
    if jitter is not None or jacdiag or os.environ.get('HMCX_RMHMC_FORCE_CTA') == '1':
        return False
    if isinstance(target, T.GaussianFull):
        return True
    return isinstance(target, (T.GaussianIso, T.GaussianDiag)) and target.dim > 16


def rmhmc_run(target, params_init, num_samples, num_steps_per_sample, step_size, burn=0, jitter=None,
              softabs_const=None, explicit_binding_const=100, fixed_point_threshold=1e-5,
              fixed_point_max_iterations=1000, jitter_max_tries=10, explicit=True, softabs=False, jacdiag=False, seed=0,
              chain_offset=0, normals=None, log_uniforms=None, uniforms=None, record_ham=False, device=None):
    """The reference's sample() loop for sampler=RMHMC over C chains (one thread per chain).

    Injected (parity) mode: ``normals`` (S, C, D), ``log_uniforms`` (S, C) and -- when ``jitter`` is not None --
    ``uniforms`` (S, C, J, D): the ``torch.rand(D)`` jitter draws of every ``fisher`` call of iteration n in the
    reference's order (J = 8*L+3 for the explicit integrator)."""
    N.require_cuda()
    lib = N.load_library()
    if device is None:
        device = params_init.device if params_init.is_cuda else torch.device('cuda', torch.cuda.current_device())
    device = torch.device(device)
    nt = native_target(target, device)
    D, ld = nt.dim, N.padded_ld(nt.dim)
    S, L, burn = int(num_samples), int(num_steps_per_sample), int(burn)
    q_init = _as_rows(params_init, ld, device)
    Cn = q_init.shape[0]
    q_cur = q_init.clone()
    eps = _eps_vector(step_size, Cn, device)
    samples = torch.empty((Cn, S - burn, ld), dtype=torch.float32, device=device)
    accepted = torch.empty((Cn, S), dtype=torch.uint8, device=device)
    diverged = torch.empty((Cn, S), dtype=torch.uint8, device=device)
    ham = torch.empty((Cn, S, 2), dtype=torch.float32, device=device) if record_ham else None
    num_rejected = torch.zeros(Cn, dtype=torch.int32, device=device)
    if softabs and softabs_const is None:
        raise RuntimeError('Metric.SOFTABS needs softabs_const')

    cfg = _rmhmc_cfg(D, step_size, jitter, softabs_const, explicit_binding_const, fixed_point_threshold,
                     fixed_point_max_iterations, jitter_max_tries, explicit, softabs, jacdiag)

    rng = N.RngStruct()
    keep_alive = []
    if normals is not None:
        z = normals.detach().to(device=device, dtype=torch.float32)
        if z.dim() == 2:
            z = z.unsqueeze(1)
        if tuple(z.shape) != (S, Cn, D):
            raise RuntimeError('normals must be (S, C, D) = (%d, %d, %d), got %s' % (S, Cn, D, tuple(z.shape)))
        z = N.pad_rows(z.contiguous(), ld)
        lu = log_uniforms.detach().to(device=device, dtype=torch.float32).reshape(S, Cn).contiguous()
        rng.mode, rng.normals, rng.log_uniforms = N.RNG_INJECTED, z.data_ptr(), lu.data_ptr()
        keep_alive += [z, lu]
        if jitter is not None:
            if uniforms is None:
                raise RuntimeError('injected RMHMC with jitter needs the uniforms stream')
            u = uniforms.detach().to(device=device, dtype=torch.float32)
            if u.dim() == 3:
                u = u.unsqueeze(1)
            if u.dim() != 4 or tuple(u.shape[:2]) != (S, Cn) or u.shape[3] != D:
                raise RuntimeError('uniforms must be (S, C, J, D) = (%d, %d, J, %d), got %s'
                                   % (S, Cn, D, tuple(u.shape)))
            u = N.pad_rows(u.contiguous(), ld)
            rng.uniforms, rng.uniforms_per_iter = u.data_ptr(), u.shape[2]
            keep_alive.append(u)
    else:
        rng.mode, rng.seed, rng.chain_offset = N.RNG_PHILOX, int(seed), int(chain_offset)
    with torch.cuda.device(device):
        if _rmhmc_is_dense(nt.target, jitter, jacdiag):
            ginv_d, lower_d, log_det = const_metric_device(nt.target, softabs, softabs_const, device)
            gm = N.ConstMetricStruct()
            gm.metric_inv, gm.metric_chol, gm.log_det = ginv_d.data_ptr(), lower_d.data_ptr(), log_det
            ws = torch.empty(lib.hmcx_rmhmc_dense_workspace_bytes(Cn, D) // 4, dtype=torch.float32, device=device)
            keep_alive += [ginv_d, lower_d, ws]
            rc = lib.hmcx_rmhmc_dense_run(nt.ref(), C.byref(cfg), C.byref(gm), C.byref(rng), N.ptr(q_init),
                                          N.ptr(q_cur), N.ptr(eps), Cn, ld, L, S, burn, 0, S, N.ptr(samples),
                                          N.ptr(accepted), N.ptr(diverged), N.ptr(ham), N.ptr(num_rejected),
                                          N.ptr(ws), N.stream_ptr(device))
            N.check(rc, 'hmcx_rmhmc_dense_run')
        else:
            rc = lib.hmcx_rmhmc_run(nt.ref(), C.byref(cfg), C.byref(rng), N.ptr(q_init), N.ptr(q_cur), N.ptr(eps), Cn,
                                    ld, L, S, burn, 0, S, N.ptr(samples), N.ptr(accepted), N.ptr(diverged), N.ptr(ham),
                                    N.ptr(num_rejected), N.stream_ptr(device))
            N.check(rc, 'hmcx_rmhmc_run')
    res = HMCResult(samples, accepted, diverged, ham, eps, num_rejected, D, S)
    res.eps_trace = None
    res.final_state = q_cur[:, :D]
    res._keep_alive = keep_alive
    return res


def _rmhmc_cfg(D, step_size, jitter, softabs_const, explicit_binding_const, fixed_point_threshold,
               fixed_point_max_iterations, jitter_max_tries, explicit, softabs, jacdiag):
    cfg = N.RmhmcStruct()
    cfg.integrator = 1 if explicit else 2
    cfg.metric = 3 if jacdiag else (2 if softabs else 1)
    cfg.softabs_const = float(softabs_const) if softabs_const is not None else 0.0
    cfg.jitter = float(jitter) if jitter is not None else -1.0
    cfg.pi_term = float(D * torch.log(2. * torch.tensor(math.pi)))                              # samplers.py:711-712
    eps0 = float(step_size) if not torch.is_tensor(step_size) else float(step_size.reshape(-1)[0])
    cfg.cos_2we = float(torch.cos(torch.FloatTensor([2 * explicit_binding_const * eps0])))    # :435
    cfg.sin_2we = float(torch.sin(torch.FloatTensor([2 * explicit_binding_const * eps0])))    # :436
    cfg.fixed_point_threshold = float(fixed_point_threshold)
    cfg.fixed_point_max_iterations = int(fixed_point_max_iterations)
    cfg.jitter_max_tries = int(jitter_max_tries)
    return cfg


def _standalone_rng(jitter, uniforms, seed, Cn, D, ld, device, keep):
    """Jitter draws of a stand-alone RMHMC call: injected ``uniforms`` (C, J, D) / (J, D) in fisher() call order, else
    Philox keyed by ``seed``."""
    rng = N.RngStruct()
    if jitter is not None and uniforms is not None:
        u = uniforms.detach().to(device=device, dtype=torch.float32)
        if u.dim() == 2:
            u = u.unsqueeze(0)
        if u.dim() != 3 or u.shape[0] != Cn or u.shape[2] != D:
            raise RuntimeError('uniforms must be (C, J, D) = (%d, J, %d), got %s' % (Cn, D, tuple(u.shape)))
        u = N.pad_rows(u.contiguous(), ld)
        rng.mode, rng.uniforms, rng.uniforms_per_iter = N.RNG_INJECTED, u.data_ptr(), u.shape[1]
        keep.append(u)
    else:
        rng.mode, rng.seed = N.RNG_PHILOX, int(seed)
    return rng


def rmhmc_leapfrog(target, q, p, steps, step_size, jitter=None, softabs_const=None, explicit_binding_const=100,
                   fixed_point_threshold=1e-20, fixed_point_max_iterations=6, jitter_max_tries=10, explicit=True,
                   softabs=False, jacdiag=False, uniforms=None, seed=0, device=None):
    """Batched samplers.leapfrog with sampler=RMHMC (explicit :389-462 / implicit :305-387), D <= 64.  q, p: (C, D) or
    (D,).  Returns (q_traj (L, C, D), p_traj (L, C, D), q_copy (C, D), p_copy (C, D), failed (C,) uint8)."""
    N.require_cuda()
    lib = N.load_library()
    device = torch.device(device if device is not None else (q.device if q.is_cuda else 'cuda'))
    nt = native_target(target, device)
    D, ld = nt.dim, N.padded_ld(nt.dim)
    qd, pd = _as_rows(q, ld, device), _as_rows(p, ld, device)
    Cn, L = qd.shape[0], int(steps)
    eps = _eps_vector(step_size, Cn, device)
    cfg = _rmhmc_cfg(D, step_size, jitter, softabs_const, explicit_binding_const, fixed_point_threshold,
                     fixed_point_max_iterations, jitter_max_tries, explicit, softabs, jacdiag)
    keep = []
    rng = _standalone_rng(jitter, uniforms, seed, Cn, D, ld, device, keep)
    q_traj = torch.zeros((L, Cn, ld), dtype=torch.float32, device=device)
    p_traj = torch.zeros_like(q_traj)
    q_copy = torch.zeros((Cn, ld), dtype=torch.float32, device=device)
    p_copy = torch.zeros_like(q_copy)
    failed = torch.zeros(Cn, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        rc = lib.hmcx_rmhmc_leapfrog(nt.ref(), C.byref(cfg), C.byref(rng), N.ptr(qd), N.ptr(pd), N.ptr(eps), Cn, ld, L,
                                     N.ptr(q_traj), N.ptr(p_traj), N.ptr(q_copy), N.ptr(p_copy), N.ptr(failed),
                                     N.stream_ptr(device))
    N.check(rc, 'hmcx_rmhmc_leapfrog')
    torch.cuda.current_stream(device).synchronize()          # `keep` / the structs may go out of scope now
    return q_traj[..., :D], p_traj[..., :D], q_copy[:, :D], p_copy[:, :D], failed


def rmhmc_hamiltonian(target, q, p, jitter=None, softabs_const=None, softabs=False, jacdiag=False, uniforms=None, seed=0,
                      device=None):
    """Batched rm_hamiltonian (samplers.py:677-736): (H (C,), failed (C,) uint8)."""
    N.require_cuda()
    lib = N.load_library()
    device = torch.device(device if device is not None else (q.device if q.is_cuda else 'cuda'))
    nt = native_target(target, device)
    D, ld = nt.dim, N.padded_ld(nt.dim)
    qd, pd = _as_rows(q, ld, device), _as_rows(p, ld, device)
    Cn = qd.shape[0]
    cfg = _rmhmc_cfg(D, 0.0, jitter, softabs_const, 100, 1e-5, 1, 10, True, softabs, jacdiag)
    keep = []
    rng = _standalone_rng(jitter, uniforms, seed, Cn, D, ld, device, keep)
    H = torch.zeros(Cn, dtype=torch.float32, device=device)
    failed = torch.zeros(Cn, dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        rc = lib.hmcx_rmhmc_hamiltonian(nt.ref(), C.byref(cfg), C.byref(rng), N.ptr(qd), N.ptr(pd), Cn, ld, N.ptr(H),
                                        N.ptr(failed), N.stream_ptr(device))
    N.check(rc, 'hmcx_rmhmc_hamiltonian')
    torch.cuda.current_stream(device).synchronize()
    return H, failed


def gemm_nt(A, B):
    """D = A @ B.T on tcgen05 tensor cores with 3xTF32 split operands (fp32-accurate).  A (M,K), B (N,K) fp32 CUDA;
    M, N multiples of 128, K a multiple of 32."""
    N.require_cuda()
    lib = N.load_library()
    A = A.detach().to(torch.float32).contiguous()
    B = B.detach().to(device=A.device, dtype=torch.float32).contiguous()
    D = torch.empty((A.shape[0], B.shape[0]), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        rc = lib.hmcx_gemm_nt_tf32x3(N.ptr(A), N.ptr(B), N.ptr(D), A.shape[0], B.shape[0], A.shape[1],
                                     N.stream_ptr(A.device))
    N.check(rc, 'hmcx_gemm_nt_tf32x3')
    return D

