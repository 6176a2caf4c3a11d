"""C5 at the BASELINE size: norm-wise error of the fp32 kernels over the WHOLE 16 x 4096 x 4096 field against the C oracle
(double precision), for the raw rsqrt.approx and (PSAD_NVRTC_EXTRA=-DPSAD_RSQRT_NEWTON=1) the Newton-refined one.

    python scripts/c5_fullsize_accuracy.py [seeds...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle.cgen import compile_c
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import make_config

shape = (16, 4096, 4096)
op = make_config('c5', shape=shape, boundary_handling='zeros')
res = {'extra': os.environ.get('PSAD_NVRTC_EXTRA', '')}
for seed in [int(a) for a in sys.argv[1:]] or [77, 1234]:
    g = torch.Generator(device='cuda')
    g.manual_seed(seed)
    arrs = {'u': torch.rand(shape, generator=g, device='cuda') * 0.9 + 0.1, 'f': torch.rand(shape, generator=g, device='cuda') * 0.9 + 0.1,
            'diffg': torch.randn(shape, generator=g, device='cuda')}
    for name in ('g', 'diffu', 'difff'):
        arrs[name] = torch.empty(shape, device='cuda')
    for assigns, ir, tag in ((op.forward_assignments, op.forward_ast_gpu, 'forward'), (op.backward_assignments, op.backward_ast_gpu, 'adjoint')):
        k = CompiledKernel(ir)
        k(**{f.name: arrs[f.name] for f in k.fields})
        torch.cuda.synchronize()
        ck = compile_c(assigns, 'zeros', 'tv_%s_full' % tag, 'strict')
        outs = {f.name for f in ir.output_fields}
        host = {f.name: (np.zeros(shape, np.float32) if f.name in outs else arrs[f.name].cpu().numpy()) for f in ir.all_fields}
        ck(**{n: host[n] for n in ck.field_names})
        for n in sorted(outs):
            got = arrs[n].cpu().numpy().astype(np.float64)
            err = np.abs(got - host[n])
            res['seed %d %s %s' % (seed, tag, n)] = {'max_rel': float(err.max() / np.abs(host[n]).max()), 'max_ref': float(np.abs(host[n]).max()),
                                                      'rms_rel': float(np.sqrt((err ** 2).mean()) / np.abs(host[n]).max()),
                                                      'cells_above_5e-7': int((err > 5e-7 * np.abs(host[n]).max()).sum())}
print(json.dumps(res, indent=1))
