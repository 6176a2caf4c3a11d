"""Time loop on slabs, fused pairs of steps with one two-plane halo exchange per pair == unsharded fused launches
(run under torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        scripts/check_slab_steps.py [c3|c4] [zeros|none] [steps]
Every rank also runs the whole (small) global field on its own GPU with ``CompiledKernel.run_steps(fuse=True)`` and
compares its slab with the matching planes: bit for bit.  Single steps (``fuse=False``) are compared the same way, and
so is a chain of two steps through the slab autograd Function (``create_slab_autograd_function``), outputs and gradients.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import make_config
from pystencils_autodiff_b200.datahandling import SlabDataHandling

SHAPES = {'c3': (24, 40, 256), 'c4': (16, 24, 128)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
    bh = None if (len(sys.argv) > 2 and sys.argv[2] == 'none') else 'zeros'
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    local = SHAPES[name]
    gshape = (local[0] * world,) + local[1:]
    op_g = make_config(name, shape=gshape, boundary_handling=bh)
    ir_g = op_g.forward_ast_gpu
    halo = max(ir_g.halo(ir_g.input_fields[0].name)[0])
    dtype = ir_g.input_fields[0].dtype.numpy_dtype
    g = torch.Generator(device='cpu')
    g.manual_seed(7)
    glob = torch.randn(gshape, generator=g, dtype=torch.float64).to(getattr(torch, str(dtype))).to(dev)
    sl = slice(rank * local[0], (rank + 1) * local[0])
    kg = CompiledKernel(ir_g)
    ok = True
    for fuse in (False, True):
        dh = SlabDataHandling(gshape, rank, world, 2 * halo, dev)
        dh.add_arrays('u, out', dtype=dtype)
        op_l = make_config(name, shape=dh.dec.local_shape, boundary_handling=bh)
        kl = CompiledKernel(op_l.forward_ast_gpu)
        dh.owned('u').copy_(glob[sl])
        res = dh.run_steps(kl, steps, fuse=fuse)
        ref = kg.run_steps(glob, steps, fuse=fuse)
        torch.cuda.synchronize()
        same = torch.equal(res[dh.dec.owned], ref[sl])
        ok = ok and same
        print('[rank %d] %s %s steps=%d fuse=%s: %s' % (rank, name, bh, steps, fuse, 'IDENTICAL' if same else
              'DIFFERENT (max %.3e)' % float((res[dh.dec.owned] - ref[sl]).abs().max())), flush=True)
    # autograd: two chained steps through the slab Function == the unsharded Function on the global field
    from pystencils_autodiff_b200.datahandling import create_slab_autograd_function
    dh = SlabDataHandling(gshape, rank, world, halo, dev)
    Step = create_slab_autograd_function(make_config(name, shape=local, boundary_handling=bh), dh)
    r = torch.randn(gshape, generator=g, dtype=torch.float64).to(glob.dtype).to(dev)
    u = glob[sl].clone().requires_grad_(True)
    (o1,) = Step.apply(u)
    (o2,) = Step.apply(o1)
    (o2 * r[sl]).sum().backward()
    Whole = op_g.create_tensorflow_op(backend='torch_native', use_cuda=True)
    ug = glob.clone().requires_grad_(True)
    (g1,) = Whole.apply(ug)
    (g2,) = Whole.apply(g1)
    (g2 * r).sum().backward()
    torch.cuda.synchronize()
    same = torch.equal(o2, g2[sl]) and torch.equal(u.grad, ug.grad[sl])
    ok = ok and same
    print('[rank %d] %s %s autograd chain of 2 slab steps: %s' % (rank, name, bh, 'IDENTICAL' if same else 'DIFFERENT'), flush=True)
    # the same chain as ONE Function: `steps` unrolled steps on the slab == create_unrolled_torch_op on the global field
    # (pairs fused where the unsharded op fuses them by default: 4-byte fields)
    from pystencils_autodiff_b200.datahandling import create_slab_unrolled_function
    fuse = glob.element_size() == 4
    dh2 = SlabDataHandling(gshape, rank, world, (2 if fuse else 1) * halo, dev)
    Many = create_slab_unrolled_function(make_config(name, shape=local, boundary_handling=bh), dh2, steps, fuse=fuse)
    WholeMany = op_g.create_unrolled_torch_op(steps, fuse=fuse)     # like with like: a fused pair rounds differently
    u = glob[sl].clone().requires_grad_(True)
    (o,) = Many.apply(u)
    (o * r[sl]).sum().backward()
    ug = glob.clone().requires_grad_(True)
    (og,) = WholeMany.apply(ug)
    (og * r).sum().backward()
    torch.cuda.synchronize()
    same = torch.equal(o, og[sl]) and torch.equal(u.grad, ug.grad[sl])
    ok = ok and same
    print('[rank %d] %s %s %d unrolled slab steps as one Function (fuse=%s): %s'
          % (rank, name, bh, steps, fuse, 'IDENTICAL' if same else 'DIFFERENT'), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
