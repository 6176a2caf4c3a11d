"""Sharded vs unsharded equality of forward + adjoint (run under torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/check_slab.py [c3|c4|c2|c5] [zeros|none]
Every rank also evaluates the whole (small) global field on its own GPU and compares its slab with the matching
planes: the two must agree bit for bit (same kernels, same per-cell arithmetic).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import make_config
from pystencils_autodiff_b200.datahandling import SlabStencilOp

SHAPES = {'c2': (96, 256), 'c3': (24, 40, 256), 'c4': (16, 24, 128), 'c5': (4, 48, 128)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
    bh = None if (len(sys.argv) > 2 and sys.argv[2] == 'none') else 'zeros'
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    local = SHAPES[name]
    gshape = (local[0] * world,) + local[1:]
    op_local = make_config(name, shape=local, boundary_handling=bh)
    slab = SlabStencilOp(op_local, local, rank, world, dev)
    # identical global data on every rank
    g = torch.Generator(device='cpu')
    g.manual_seed(7)
    glob = {}
    for f in op_local.forward_input_fields:
        glob[f.name] = (torch.rand(gshape, generator=g, dtype=torch.float64) * 0.9 + 0.1).to(slab.dh.gpu_arrays[f.name].dtype)
    for f in op_local.backward_input_fields:
        if f not in op_local.forward_input_fields:
            glob[f.name] = torch.randn(gshape, generator=g, dtype=torch.float64).to(slab.dh.gpu_arrays[f.name].dtype)
    sl = slice(rank * local[0], (rank + 1) * local[0])
    for n, t in glob.items():
        slab.dh.owned(n).copy_(t[sl].to(dev))
    slab.forward()
    slab.backward()
    torch.cuda.synchronize()
    # unsharded reference on this GPU
    op_g = make_config(name, shape=gshape, boundary_handling=bh)
    fk, bk = CompiledKernel(op_g.forward_ast_gpu), CompiledKernel(op_g.backward_ast_gpu)
    arrays = {n: t.to(dev) for n, t in glob.items()}
    for k in (fk, bk):
        for f in k.ir.output_fields:
            arrays[f.name] = torch.empty(gshape, dtype=slab.dh.gpu_arrays[f.name].dtype, device=dev)
        k(**{f.name: arrays[f.name] for f in k.fields})
    torch.cuda.synchronize()
    ok = True
    for f in list(op_local.forward_output_fields) + list(op_local.backward_output_fields):
        a = slab.dh.owned(f.name)
        b = arrays[f.name][sl]
        same = torch.equal(a, b)
        err = (a.double() - b.double()).abs().max().item()
        print('[rank %d] %s %s %s: %s (max abs diff %.3e) variants %s' % (rank, name, bh, f.name, 'IDENTICAL' if same else 'DIFFERENT',
                                                                     err, slab.variants()), flush=True)
        ok &= same
    res = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if res.item() == 1 else 1)


if __name__ == '__main__':
    main()
