"""Reproduce bench.py's `fused_steps.time_loop_api` leg for one workload with full tracebacks (debug aid).
    python scripts/debug_c2_timeloop.py [c2|c3] [order: single,fused | fused,single]"""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pystencils_autodiff_b200 import runtime
from pystencils_autodiff_b200.configs import CONFIG_SHAPES, make_config
from pystencils_autodiff_b200.datahandling import SlabStencilOp

name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
order = (sys.argv[2] if len(sys.argv) > 2 else 'single,fused').split(',')
shape = tuple(CONFIG_SHAPES[name]['shape'])
op = make_config(name, shape=shape)
slab = SlabStencilOp(op, shape, 0, 1, device=torch.device('cuda', 0))
gen = torch.Generator(device='cuda')
gen.manual_seed(0)
slab.randomize(gen)
fk = slab.fwd
fin, fout = op.forward_ast_gpu.input_fields[0].name, op.forward_ast_gpu.output_fields[0].name
print('dec.g', slab.dh.dec.g, 'arrays', {n: tuple(t.shape) for n, t in slab.dh.gpu_arrays.items()}, flush=True)
T = 8
for which in order:
    fuse = False if which == 'single' else None
    try:
        tl = slab.dh.create_timeloop(use_cuda_graph=True, fuse_steps=fuse)
        tl.add_call(fk, {})
        tl.swap(fin, fout)
        n0 = runtime.launch_count()
        tl.run(T)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        tl.run(T)
        tl.run(T)
        b.record()
        torch.cuda.synchronize()
        print(which, 'ok: ms per time step', a.elapsed_time(b) / (2 * T), 'fused', tl.fused_last_run, 'launches', runtime.launch_count() - n0,
              'graphs', len(tl._graphs), flush=True)
    except Exception:
        print(which, 'FAILED', flush=True)
        traceback.print_exc()
        sys.stderr.flush()
        break
