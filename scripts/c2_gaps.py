"""Where do C2's ~14 us per kernel between the ncu kernel duration (82 us) and the event-timed step go?  Times K steps of
forward + adjoint (a) as plain stream launches, (b) as ONE captured CUDA graph of K steps, (c) as K replays of a one-step graph.

    python scripts/c2_gaps.py [c2|c5|c3]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import CONFIG_SHAPES, make_config

name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
shape = CONFIG_SHAPES[name]['shape']
op = make_config(name, shape=shape, boundary_handling='zeros')
fk, bk = CompiledKernel(op.forward_ast_gpu), CompiledKernel(op.backward_ast_gpu)
dt = getattr(torch, str(fk.fields[0].dtype.numpy_dtype))
arrs = {}
for k in (fk, bk):
    for f in k.fields:
        arrs.setdefault(f.name, torch.rand(shape, dtype=dt, device='cuda') + 0.1)
K = 20


def step():
    fk(**{f.name: arrs[f.name] for f in fk.fields})
    bk(**{f.name: arrs[f.name] for f in bk.fields})


def timed(fn, reps=5):
    best = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best.append(a.elapsed_time(b) / K)
    best.sort()
    return {'median_ms_per_step': best[len(best) // 2], 'best_ms_per_step': best[0]}


for _ in range(5):
    step()
res = {'workload': name, 'steps': K}
res['stream_launches'] = timed(lambda: [step() for _ in range(K)])
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    gK, g1 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(gK, stream=side):
        for _ in range(K):
            step()
    with torch.cuda.graph(g1, stream=side):
        step()
torch.cuda.current_stream().wait_stream(side)
res['one_graph_of_K_steps'] = timed(gK.replay)
res['K_replays_of_one_step_graph'] = timed(lambda: [g1.replay() for _ in range(K)])
print(json.dumps(res, indent=1))
