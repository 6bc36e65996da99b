import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ['x']
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import time_flow as tf
for kind, C, D, S, L in [('rmhmc_explicit', 8192, 64, 50, 10), ('fullmass_hmc', 4096, 128, 50, 10), ('rmhmc_explicit', 2048, 64, 100, 10), ('rmhmc_explicit', 1024, 64, 100, 10)]:
    for w in ('4', '8'):
        os.environ['HMCX_FLOW_W'] = w
        os.environ['HMCX_FLOW_SMALL'] = '1'
        ms, r, mv = tf.run(kind, C, D, S, L)
        print(kind, C, D, 'w', w, '%.3f ms  %.3g chain-steps/s' % (ms, C * S * L / ms * 1e3), flush=True)
