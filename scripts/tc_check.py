import sys; sys.path.insert(0,".")
import torch, json
import hamiltorch_b200 as hb
from hamiltorch_b200 import engine, targets as T
from oracle import cases
from scripts.bench_paths import cfg4
model, x, y = cases.mlp_problem(seed=7, n=1024, n_in=64, hidden=128)
def mk(tc):
    d=[T.MLPTarget.from_model(model, x[m*256:(m+1)*256], y[m*256:(m+1)*256], None, 100., prior_scale=4) for m in range(4)]
    d[0].tensor_cores = tc
    return d
q = hb.util.flatten(model).detach()[None] + 0.05*torch.randn(3, 8449)
g1,l1 = engine.grad_log_prob(mk(0), q, split=1)
g0,l0 = engine.grad_log_prob(mk(1), q, split=1)
print("max diff tc vs simt", (g1-g0).abs().max().item(), "scale", g0.abs().max().item(), "identical", torch.equal(g1,g0), l1.tolist(), l0.tolist())
print(json.dumps(cfg4()))
print(json.dumps(cfg4(C=8)))
print(json.dumps(cfg4(C=148)))
