"""Periodic domains on GPUs: time loops on slabs that wrap around along the decomposed axis (one process, or one rank per
GPU under torchrun).

    python scripts/check_periodic.py [c3|c4] [steps]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 \
        scripts/check_periodic.py c3 5
Every rank also runs the whole (small) global field as ONE periodic rank on its own GPU (ghost planes by device copies) and
compares its slab with the matching planes bit for bit — single steps and fused pairs — and that one-rank result with a
plain torch restatement of the periodic stencil (``torch.roll`` along dim 0, zero padding along the other axes).  With two
ranks both neighbours are the same peer: the grouped ncclSend/ncclRecv of ``psad_halo_exchange`` must pair up in issue order."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import make_config
from pystencils_autodiff_b200.datahandling import SlabDataHandling

SHAPES = {'c3': (24, 40, 256), 'c4': (16, 24, 128)}


def torch_periodic_step(name, u):
    """One step of C3 / C4 (configs.py) on a field periodic along dim 0 and zero outside along dims 1, 2, in float64."""
    import itertools
    import torch.nn.functional as F
    u = u.double()
    p = F.pad(u, (1, 1, 1, 1))                         # zeros along y and x
    ny, nx = u.shape[1], u.shape[2]

    def sh(dz, dy, dx):
        return torch.roll(p, -dz, 0)[:, 1 + dy:1 + dy + ny, 1 + dx:1 + dx + nx]
    if name == 'c3':
        nb = sum(sh(*o) for o in [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)])
        return u + 0.1 * (nb - 6 * u)
    w = (0.4, 0.05, 0.02, 0.0075)
    return sum(w[sum(abs(v) for v in o)] * sh(*o) for o in itertools.product((-1, 0, 1), repeat=3))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    local = SHAPES[name]
    gshape = (local[0] * world,) + local[1:]
    probe = make_config(name, shape=gshape).forward_ast_gpu
    dtype = probe.input_fields[0].dtype.numpy_dtype
    tdtype = getattr(torch, str(dtype))
    gen = torch.Generator(device='cpu')
    gen.manual_seed(17)
    glob = torch.randn(gshape, generator=gen, dtype=torch.float64).to(tdtype).to(dev)
    sl = slice(rank * local[0], (rank + 1) * local[0])
    ok = True
    for fuse in (False, None):
        results = []
        for r, w in ((rank, world), (0, 1)):          # this rank's slab, then the whole field as one periodic rank
            dh = SlabDataHandling(gshape, r, w, 2, dev, periodic=True)
            dh.add_arrays('u, out', dtype=dtype)
            kernel = CompiledKernel(make_config(name, shape=dh.dec.local_shape).forward_ast_gpu)
            dh.owned('u').copy_(glob[sl] if w > 1 else glob)
            before = len(dh.call_queue)
            dh.run_steps(kernel, steps, fuse=fuse)
            torch.cuda.synchronize()
            results.append(dh.owned('u').clone())
            fused = sum(1 for c in dh.call_queue[before:] if c[0] == 'KernelCall' and len(c) > 2)
            assert (fused > 0) == (fuse is None and steps >= 2), (fuse, dh.call_queue[before:])
            dh.close()
        same = torch.equal(results[0], results[1][sl] if world > 1 else results[1])
        ref = glob
        for _ in range(steps):
            ref = torch_periodic_step(name, ref)
        err = float((results[1].double() - ref).abs().max() / ref.abs().max())
        tol = 2e-6 if dtype.itemsize == 4 else 1e-13
        ok = ok and same and err <= tol
        print('[rank %d] %s periodic steps=%d %s: %s, one periodic rank vs torch.roll restatement %.2e (tol %.0e)'
              % (rank, name, steps, 'fused pairs' if fuse is None else 'single steps',
                 'IDENTICAL' if same else 'DIFFERENT', err, tol), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
