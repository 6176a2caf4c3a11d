"""Separate vs fused forward+adjoint timing: python scripts/fused_bench.py c1 8192 8192"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pystencils_autodiff_b200.configs import make_config, CONFIG_SHAPES
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel, numpy_dtype_to_torch

name = sys.argv[1]
shape = tuple(int(s) for s in sys.argv[2:]) or CONFIG_SHAPES[name]['shape']
op = make_config(name, shape=shape)
cells = 1
for s in shape:
    cells *= s
ks = {'forward': CompiledKernel(op.forward_ast_gpu), 'adjoint': CompiledKernel(op.backward_ast_gpu), 'fused': op.fused_kernel_gpu}
tens = {}
for k in ks.values():
    for f in k.fields:
        tens.setdefault(f.name, torch.rand(shape, dtype=numpy_dtype_to_torch(f.dtype.numpy_dtype), device='cuda') + 0.5)
res = {}
for nm, k in ks.items():
    args = {f.name: tens[f.name] for f in k.fields}
    for _ in range(5):
        k(**args)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20):
        k(**args)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    res[nm] = ms
    print('%-8s %-8s %s %.3f ms  %d B/cell  %.0f GB/s  variant=%s' % (name, nm, shape, ms, k.ir.bytes_per_cell(),
                                                                cells * k.ir.bytes_per_cell() / ms / 1e6, k.last_variant))
print('pair separate %.3f ms, fused %.3f ms (%.2fx)' % (res['forward'] + res['adjoint'], res['fused'],
                                                      (res['forward'] + res['adjoint']) / res['fused']))
