import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pystencils_autodiff_b200.configs import make_config
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
op = make_config('c2', shape=(64, 128))
k = CompiledKernel(op.forward_ast_gpu)
u = torch.randn(64, 128, device='cuda'); out = torch.empty_like(u)
for _ in range(10): k(u=u, out=out)
torch.cuda.synchronize()
n = 2000
t0 = time.perf_counter()
for _ in range(n): k(u=u, out=out)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('host time per launch %.1f us; incl. drain %.1f us' % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
ug = u.clone().requires_grad_(True)
for _ in range(10):
    (o,) = fn.apply(ug); o.backward(out)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(500):
    (o,) = fn.apply(ug); o.backward(out)
torch.cuda.synchronize(); print('Function.apply + backward per step %.1f us' % ((time.perf_counter() - t0) / 500 * 1e6))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(500): k(u=u, out=out)
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
