"""Small launches of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python scripts/sanitize_small.py

The row-exchange variant of the fused-pair kernel passes intermediate rows between warps through shared memory with ONE named
barrier per plane (emit_chain.py): racecheck is the tool that would see a missing / misplaced barrier there."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import make_config

torch.manual_seed(0)
for name, shape in (('c3', (10, 60, 256)), ('c4', (8, 40, 128)), ('c2', (96, 256)), ('c5', (2, 48, 256)), ('c1', (20, 30))):
    for bh in ('zeros', None):
        op = make_config(name, shape=shape, boundary_handling=bh)
        for ir in (op.forward_ast_gpu, op.backward_ast_gpu):
            k = CompiledKernel(ir)
            dt = getattr(torch, str(k.fields[0].dtype.numpy_dtype))
            arrs = {f.name: torch.rand(shape, dtype=dt, device='cuda') + 0.5 for f in k.fields}
            k(**arrs)
            variants = [k.last_instance]
            if k.fused_steps_reason() is None and len(k.fields) == 2:
                k(**arrs, _variant='march_x2')
                variants.append('march_x2 (%s)' % k.emitted('march_x2').name)
            torch.cuda.synchronize()
            print('%-28s %-6s %s' % (ir.function_name, bh, variants), flush=True)
print('done')
