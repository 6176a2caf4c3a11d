"""Peer halos (``SlabDataHandling(peer_halo=True)``: ghost planes read by the kernel from the neighbouring GPUs' arrays over
NVLink, one launch per kernel, no exchange) against the unsharded kernels and against the NCCL path (run under torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/check_peer_halo.py [c3|c4] [zeros|none] [steps] [--time]
Correctness: a time loop of ``steps`` steps (single steps, then fused pairs) on small slabs, every rank compares its planes
with the unsharded run on its own GPU — bit for bit; many more steps than ranks, so a missed wait shows up as a difference.
``--time``: the full-size workload, ms per step for NCCL exchange vs peer halos (forward+adjoint pair and the time loop).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import CONFIG_SHAPES, make_config
from pystencils_autodiff_b200.datahandling import SlabDataHandling, SlabStencilOp

SHAPES = {'c3': (24, 40, 256), 'c4': (16, 24, 128)}


def correctness(name, bh, steps, rank, world, dev):
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
        dh = SlabDataHandling(gshape, rank, world, 2 * halo, dev, peer_halo=True)
        dh.add_arrays('u, out', dtype=dtype)
        op_l = make_config(name, shape=dh.dec.local_shape, boundary_handling=bh)
        kl = CompiledKernel(op_l.forward_ast_gpu)
        dh.owned('u').copy_(glob[sl])
        dh.peer.dirty = True
        res = dh.run_steps(kl, steps, fuse=fuse)
        ref = kg.run_steps(glob, steps, fuse=fuse)
        torch.cuda.synchronize()
        same = torch.equal(res[dh.dec.owned], ref[sl]) and dh.peer.errors() == 0
        ok = ok and same
        print('[rank %d] %s %s steps=%d fuse=%s peer halos (%s, %d launches): %s' % (
            rank, name, bh, steps, fuse, kl.last_instance, dh.peer.seq, 'IDENTICAL' if same else
            'DIFFERENT (max %.3e, errors %d)' % (float((res[dh.dec.owned] - ref[sl]).abs().max()), dh.peer.errors())), flush=True)
        dh.close()
    return ok


def mixed_sequence(name, bh, steps, rank, world, dev):
    """Launches inside and outside the peer protocol on the same arrays: the stencil with peer halos (u -> out), a pointwise
    kernel without any reach (out -> u: overwrites planes the neighbours' stencil launch may still be reading — ordered by
    psad_peer_wait), the same stencil through the NCCL path (its march instances removed: a 'foreign' launch), pointwise
    again.  Against the same sequence unsharded."""
    import pystencils_autodiff_b200 as ps
    local = SHAPES[name]
    gshape = (local[0] * world,) + local[1:]
    op_g = make_config(name, shape=gshape, boundary_handling=bh)
    ir_g = op_g.forward_ast_gpu
    halo = max(ir_g.halo(ir_g.input_fields[0].name)[0])
    dtype = ir_g.input_fields[0].dtype.numpy_dtype
    g = torch.Generator(device='cpu')
    g.manual_seed(11)
    glob = torch.randn(gshape, generator=g, dtype=torch.float64).to(getattr(torch, str(dtype))).to(dev)
    sl = slice(rank * local[0], (rank + 1) * local[0])

    def pointwise(shape):
        u, out = ps.fields('u, out: %s[%s]' % (dtype, ','.join(str(v) for v in shape)))
        return ps.AutoDiffOp(ps.AssignmentCollection([ps.Assignment(u.center, 0.5 * out.center + 0.25)]), op_name='halve',
                             boundary_handling='zeros').forward_ast_gpu

    dh = SlabDataHandling(gshape, rank, world, halo, dev, peer_halo=True)
    dh.add_arrays('u, out', dtype=dtype)
    lshape = dh.dec.local_shape
    k_st = CompiledKernel(make_config(name, shape=lshape, boundary_handling=bh).forward_ast_gpu)
    k_foreign = CompiledKernel(make_config(name, shape=lshape, boundary_handling=bh).forward_ast_gpu)
    for v in [v for v in k_foreign._emitted if v.startswith('march')]:
        del k_foreign._emitted[v]                  # only the generic kernel is left: never a peer launch
    k_pw = CompiledKernel(pointwise(lshape))
    kg_st, kg_pw = CompiledKernel(ir_g), CompiledKernel(pointwise(gshape))
    kg_foreign = CompiledKernel(make_config(name, shape=gshape, boundary_handling=bh).forward_ast_gpu)
    for v in [v for v in kg_foreign._emitted if v.startswith('march')]:
        del kg_foreign._emitted[v]
    dh.owned('u').copy_(glob[sl])
    dh.peer.dirty = True
    u_ref, out_ref = glob.clone(), torch.zeros_like(glob)
    for i in range(steps):
        dh.run_kernel(k_st if i % 2 == 0 else k_foreign, halo_fields=['u'])
        dh.run_kernel(k_pw)
        (kg_st if i % 2 == 0 else kg_foreign)(u=u_ref, out=out_ref)
        kg_pw(u=u_ref, out=out_ref)
    torch.cuda.synchronize()
    same = torch.equal(dh.owned('u'), u_ref[sl]) and dh.peer.errors() == 0
    print('[rank %d] %s %s mixed sequence, %d steps (%s / %s / %s, %d launches counted): %s' % (
        rank, name, bh, steps, k_st.last_instance, k_foreign.last_instance, k_pw.last_instance, dh.peer.seq,
        'IDENTICAL' if same else 'DIFFERENT (max %.3e, errors %d)' % (float((dh.owned('u') - u_ref[sl]).abs().max()), dh.peer.errors())),
        flush=True)
    dh.close()
    return same


def timing(name, rank, world, dev, steps=20, reps=3):
    """ms per forward+adjoint step: every GPU alone on its slab (no neighbours, no halos), NCCL exchange, peer halos — the
    three operators are built once and timed in rotating order, each timed region started from an idle GPU (this pool's
    GPUs power-cap under sustained load: whatever is measured last would otherwise look slowest)."""
    import time
    shape = tuple(CONFIG_SHAPES[name]['shape'])
    if os.environ.get('PSAD_CHECK_SHAPE'):           # e.g. 128,1024,1024: the per-GPU slab of the N = 8 strong-scaling run
        shape = tuple(int(v) for v in os.environ['PSAD_CHECK_SHAPE'].split(','))
    out = {'workload': name, 'n_gpus': world, 'per_gpu_shape': list(shape)}
    slabs = {}
    for mode in ('alone', 'nccl', 'peer'):
        op = make_config(name, shape=shape, boundary_handling='zeros')
        if mode == 'alone':      # every GPU on its own slab without neighbours: what the kernels take with no halo at all
            slab = SlabStencilOp(op, shape, 0, 1, dev)
        else:
            slab = SlabStencilOp(op, shape, rank, world, dev, peer_halo=(mode == 'peer'))
        g = torch.Generator(device=dev)
        g.manual_seed(3 + rank)
        slab.randomize(g)
        slabs[mode] = slab

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    order = ['alone', 'nccl', 'peer']
    for rep in range(reps):
        for mode in order[rep % 3:] + order[:rep % 3]:
            slab = slabs[mode]
            for _ in range(3):
                slab.forward()
                slab.backward()
            barrier()
            time.sleep(1.5)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                slab.forward()
                slab.backward()
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out.setdefault('%s_ms_per_step' % mode, []).append(round(float(t.item()), 4))
    out['peer_errors'] = slabs['peer'].dh.peer.errors()
    out['peer_launches'] = slabs['peer'].dh.peer.seq
    out['peer_equals_nccl'] = all(torch.equal(slabs['nccl'].dh.owned(n), slabs['peer'].dh.owned(n))
                                  for n in ('out', 'diffu') if n in slabs['nccl'].dh.gpu_arrays)
    best = {m: min(out['%s_ms_per_step' % m]) for m in order}
    out['best_ms_per_step'] = best
    out['speedup'] = best['nccl'] / best['peer']
    out['overhead_us_per_step'] = {'nccl': round(1e3 * (best['nccl'] - best['alone']), 1), 'peer': round(1e3 * (best['peer'] - best['alone']), 1)}
    slabs['peer'].dh.close()
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    name = args[0] if args else 'c3'
    bh = None if (len(args) > 1 and args[1] == 'none') else 'zeros'
    steps = int(args[2]) if len(args) > 2 else 9
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    ok = correctness(name, bh, steps, rank, world, dev)
    ok = mixed_sequence(name, bh, steps, rank, world, dev) and ok
    res = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if '--time' in sys.argv and int(res.item()) == 1:
        t = timing(name, rank, world, dev)
        if rank == 0:
            print(json.dumps(t), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if res.item() == 1 else 1)


if __name__ == '__main__':
    main()
