"""Host cost of the launch path: raw CompiledKernel launches, and Function.apply + backward.

    python scripts/launch_overhead.py            # tiny arrays: host-bound by construction
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pystencils_autodiff_b200 import runtime
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import make_config

res = {}
# PSAD_NO_LAUNCH_CACHE=1 python scripts/launch_overhead.py  -> the round-1 behaviour (arguments re-validated and re-packed in
# Python, parameter block rebuilt and tensor maps re-encoded in C on every launch) for comparison
uncached = bool(os.environ.get('PSAD_NO_LAUNCH_CACHE'))


class _Forget(dict):
    def get(self, key, default=None):
        return None

    def __setitem__(self, key, value):
        pass


op = make_config('c2', shape=(64, 128))
k = CompiledKernel(op.forward_ast_gpu)
u = torch.randn(64, 128, device='cuda')
out = torch.empty_like(u)
if uncached:
    k._fast = _Forget()
for _ in range(10):
    k(u=u, out=out)
torch.cuda.synchronize()
n = 3000
t0 = time.perf_counter()
for _ in range(n):
    k(u=u, out=out)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
res['launch_cache_enabled'] = not uncached
res['raw_launch_host_us'] = (t1 - t0) / n * 1e6
res['raw_launch_incl_drain_us'] = (t2 - t0) / n * 1e6
op = make_config('c2', shape=(64, 128))
fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
u = torch.randn(64, 128, device='cuda')
out = torch.empty_like(u)
ug = u.clone().requires_grad_(True)
for _ in range(10):
    (o,) = fn.apply(ug)
    o.backward(out)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(1000):
    (o,) = fn.apply(ug)
    o.backward(out)
    ug.grad = None
torch.cuda.synchronize()
res['function_apply_backward_us_per_step'] = (time.perf_counter() - t0) / 1000 * 1e6
# the same through plain torch ops of comparable launch count (two elementwise kernels): the floor autograd itself sets
x = u.clone().requires_grad_(True)
t0 = time.perf_counter()
for _ in range(1000):
    o = x * 2.0
    o.backward(out)
    x.grad = None
torch.cuda.synchronize()
res['torch_mul_backward_us_per_step'] = (time.perf_counter() - t0) / 1000 * 1e6
res['launch_cache'] = runtime.launch_cache_stats()
print(json.dumps(res))
import cProfile
import pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    (o,) = fn.apply(ug)
    o.backward(out)
    ug.grad = None
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
