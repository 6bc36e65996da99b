"""hamiltorch_b200 -- a B200-native (sm_100a) batched-chain HMC engine behind the hamiltorch surface.

Exports the same names as the reference's ``hamiltorch/__init__.py:1-4`` plus the engine's native additions
(``targets``, ``sample_chains``).  The CUDA library is loaded lazily by the first sampling call and that call
fails loudly if libhmcx.so is missing or no GPU is present: there is no CPU fallback.
"""
__version__ = '0.1.0'

from . import targets, util
from .samplers import (sample, sample_chains, sample_model, sample_split_model, predict_model, Sampler, Integrator,
                       Metric, leapfrog, hamiltonian, gibbs, acceptance, adaptation, define_model_log_prob,
                       define_split_model_log_prob)
from .util import set_random_seed
