"""C5 (TV-denoising gradient) accuracy and speed by arithmetic mode, against the fp64 oracle:
fp32 compute with the raw rsqrt.approx (default), with a Newton-refined reciprocal root (PSAD_NVRTC_EXTRA=-DPSAD_RSQRT_NEWTON=1),
and data_type='double' (pystencils' default promotion: fp32 fields, fp64 arithmetic).

    python scripts/c5_accuracy.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import forward_backward
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import make_config


def rel(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-300))


def timed(k, arrs, n=20):
    for _ in range(3):
        k(**arrs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        k(**arrs)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


out = {}
mode = os.environ.get('C5_MODE', 'default')
kw = {'data_type': 'double'} if mode == 'double' else {}
for bh in ('zeros', None):
    for shape, lo in (((3, 40, 136), 0.0), ((2, 512, 512), 0.0), ((2, 512, 512), 0.1)):
        worst = {}
        for seed in range(4):
            rng = np.random.default_rng(seed)
            op = make_config('c5', shape=shape, boundary_handling=bh, **kw)
            ins = {f.name: rng.uniform(lo, 1.0, shape).astype(np.float32) for f in op.forward_input_fields}
            grads = {f.name: rng.standard_normal(shape).astype(np.float32) for f in op.forward_output_fields}
            fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
            xs = [torch.from_numpy(ins[f.name]).cuda().requires_grad_(True) for f in op.forward_input_fields]
            outs = fn.apply(*xs)
            gin = torch.autograd.grad(outs, xs, [torch.from_numpy(grads[f.name]).cuda() for f in op.forward_output_fields])
            ref_out, ref_din = forward_backward(op, ins, grads)
            for f, o in zip(op.forward_output_fields, outs):
                worst[f.name] = max(worst.get(f.name, 0), rel(o.detach().cpu().numpy(), ref_out[f.name]))
            for f, g in zip(op.forward_input_fields, gin):
                worst['diff' + f.name] = max(worst.get('diff' + f.name, 0), rel(g.cpu().numpy(), ref_din['diff' + f.name]))
        out['%s %s lo=%.1f' % (bh, shape, lo)] = worst
op = make_config('c5', **kw)
shape = (16, 4096, 4096)
for ir, tag in ((op.forward_ast_gpu, 'forward'), (op.backward_ast_gpu, 'adjoint')):
    k = CompiledKernel(ir)
    arrs = {f.name: torch.rand(shape, device='cuda') + 0.1 for f in k.fields}
    out['ms_' + tag] = timed(k, arrs)
    out['regs_' + tag] = k.native(k.last_instance).attributes()['num_regs']
print(json.dumps({'mode': mode, 'extra': os.environ.get('PSAD_NVRTC_EXTRA', ''), **out}, indent=1))
