"""Mirror of the parts of hamiltorch/util.py that the hot path and its callers use.

``set_random_seed`` (util.py:11-23), ``has_nan_or_inf`` / ``LogProbError`` (util.py:92-104), the flat parameter
layout helpers ``flatten`` / ``unflatten`` / ``update_model_params_in_place`` (util.py:121-141) and the
many-chains entry ``setup_chain`` / ``multi_chain`` (util.py:385-405), which here runs all chains as ONE batched
kernel launch instead of a Python loop / thread pool.
"""
import random
import time

import numpy as np
import torch


def set_random_seed(seed=None):
    """util.py:11-20: seeds python, numpy and torch (+cuda).  Unlike the reference this module does NOT reseed
    with a time seed at import (util.py:23) -- importing a library should not clobber the caller's RNG state."""
    if seed is None:
        seed = int((time.time() * 1e6) % 1e8)
    global _random_seed
    _random_seed = seed
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def has_nan_or_inf(value):
    """util.py:92-100 (including the float branch's inability to detect NaN: ``x == float('NaN')`` is False)."""
    if torch.is_tensor(value):
        value = torch.sum(value)
        return int(torch.isnan(value)) > 0 or int(torch.isinf(value)) > 0
    value = float(value)
    return (value == float('inf')) or (value == float('-inf')) or (value == float('NaN'))


class LogProbError(Exception):
    """util.py:103.  Inside the kernels a non-finite log-prob sets the chain's per-iteration ``diverged`` flag and
    the iteration is rejected (samplers.py:1045); host-side helpers raise this exception like the reference."""


# ---- flat parameter layout: model.parameters() order, each tensor row-major (util.py:121-136) -----------------
def flatten(model):
    return torch.cat([p.flatten() for p in model.parameters()])


def unflatten(model, flattened_params):
    if flattened_params.dim() != 1:
        raise ValueError('Expecting a 1d flattened_params')
    params_list = []
    i = 0
    for val in list(model.parameters()):
        length = val.nelement()
        params_list.append(flattened_params[i:i + length].view_as(val))
        i += length
    return params_list


def update_model_params_in_place(model, params):
    for weights, new_w in zip(model.parameters(), params):
        weights.data = new_w


# ---- many chains (util.py:385-405) ----------------------------------------------------------------------------
def setup_chain(sampler, prior, kwargs):
    """util.py:385-390.  The returned closure runs one chain like the reference's; it also carries what
    ``multi_chain`` needs to run all chains as one batched launch."""
    def chain(seed):
        torch.manual_seed(seed)
        params_init = prior()
        return sampler(params_init=params_init, **kwargs)
    chain._hmcx = (sampler, prior, dict(kwargs))
    return chain


def multi_chain(chain, num_workers, seeds, parallel=False):
    """util.py:392-405.  Returns list[chain] of list[sample] of (D,) tensors, like the reference.

    When ``chain`` comes from this module's ``setup_chain`` around ``hamiltorch_b200.sample`` all chains run in ONE
    persistent-kernel launch.  Each chain still consumes exactly the random stream the reference's chain(seed)
    would (``manual_seed(seed)`` -> ``prior()`` -> per iteration randn(D), rand(1)), so results match the
    reference's ``multi_chain(parallel=False)`` chain by chain; ``num_workers`` / ``parallel`` are accepted and
    ignored (the GPU is the worker pool)."""
    from . import samplers
    meta = getattr(chain, '_hmcx', None)
    if meta is None or meta[0] is not samplers.sample:
        return [chain(s) for s in seeds]
    _, prior, kw = meta
    kw = dict(kw)
    if kw.get('pass_grad') is not None:
        raise NotImplementedError('pass_grad: gradients are analytic inside the kernel')
    # the one-launch form pre-draws randn(D) + rand(1) per iteration: exactly the stream of plain HMC / HMC_NUTS with the
    # plain integrator.  SPLITTING_RAND (randperm per trajectory) and RMHMC (jitter draws) consume more: those chains run
    # one by one through sample(), which draws their stream itself.
    batched = kw.get('sampler', samplers.Sampler.HMC) in (samplers.Sampler.HMC, samplers.Sampler.HMC_NUTS) and \
        kw.get('integrator', samplers.Integrator.IMPLICIT) not in samplers._SPLIT_INTEGRATORS and \
        not isinstance(kw.get('log_prob_func'), list)
    if not batched:
        return [chain(s) for s in seeds]
    verbose = kw.pop('verbose', True)
    debug = kw.pop('debug', False)
    store_on_GPU = kw.pop('store_on_GPU', True)
    kw.pop('rng', None)
    kw.pop('seed', None)
    for dead in ('normalizing_const', 'pass_grad'):
        kw.pop(dead, None)
    log_prob_func = kw.pop('log_prob_func')
    S = kw.get('num_samples', 10)
    inits, zs, lus = [], [], []
    for s in seeds:
        torch.manual_seed(s)
        q0 = prior()
        if q0.dim() != 1:
            raise RuntimeError('params_init must be a 1d tensor.')
        z, lu = samplers._draw_reference_stream(q0.numel(), S, q0.device)
        inits.append(q0)
        zs.append(z)
        lus.append(lu)
    q0 = torch.stack(inits)
    res = samplers.sample_chains(log_prob_func, q0, rng='injected', normals=torch.stack(zs, 1),
                                 log_uniforms=torch.stack(lus, 1), **kw)
    out_dev = q0.device if store_on_GPU else torch.device('cpu')
    samples = res.samples.to(out_dev)
    nuts = kw.get('sampler', samplers.Sampler.HMC) == samplers.Sampler.HMC_NUTS
    results = []
    for c in range(len(seeds)):
        lst = list(samples[c].unbind(0))
        rate = 1 - int(res.num_rejected[c]) / S
        if verbose:
            print('Acceptance Rate {:.2f}'.format(rate))
        if debug == 2:
            results.append((lst, float(res.step_size[c]) if nuts else rate))
        else:
            results.append(lst)
    return results
