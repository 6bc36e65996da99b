"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference (AdamCobb/hamiltorch @ 19b627b) from /root/reference.  That tree only exists
in the build container, never on the GPU box: nothing under tests/ -m gpu, smoke() or bench.py may call this.
It is used by oracle/gen_golden.py (fixture generation) and by CPU tests that are skipped when the tree is
absent.

The reference needs ``termcolor`` (util.py:4, used only by eval_print) which is not installed here; a two-line
shim module is injected into sys.modules.  torch.distributions argument validation is
switched off because modern torch turns the notebook funnel's scale underflow into a ValueError (section 8a).
"""
import os
import sys
import types

REFERENCE_ROOT = '/root/reference'


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'hamiltorch'))


def import_reference():
    if not reference_available():
        raise ImportError('reference tree not present at ' + REFERENCE_ROOT)
    if 'termcolor' not in sys.modules:
        shim = types.ModuleType('termcolor')
        shim.colored = lambda s, *a, **k: s
        sys.modules['termcolor'] = shim
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch
    torch.distributions.Distribution.set_default_validate_args(False)
    import hamiltorch  # noqa: E402  (the reference package)
    if not hamiltorch.__file__.startswith(REFERENCE_ROOT):
        raise ImportError('imported a hamiltorch that is not the reference: ' + hamiltorch.__file__)
    return hamiltorch
