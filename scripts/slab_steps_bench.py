"""Time loop on slabs: T single steps (one-plane exchanges) against T/2 fused pairs (one two-plane exchange per pair).

    python scripts/slab_steps_bench.py [c3|c4] [steps]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        scripts/slab_steps_bench.py c3 8                                                  # weak scaling, workload shape per GPU
Prints one JSON line per mode (rank 0): ms per step (max over ranks, CUDA events, barrier on both sides), aggregate
Gcell-steps/s, and the largest difference between the two modes on this rank's slab."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import CONFIG_SHAPES, make_config
from pystencils_autodiff_b200.datahandling import SlabDataHandling


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    name = args[0] if args else 'c3'
    steps = int(args[1]) if len(args) > 1 else 8
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    local = tuple(CONFIG_SHAPES[name]['shape'])
    if '--strong' in sys.argv:           # the global field is the workload's shape, every rank owns 1 / world of its planes
        local = (local[0] // world,) + local[1:]
    peer = '--peer' in sys.argv          # ghost planes read by the kernels from the neighbouring GPUs (no exchange launches)
    gshape = (local[0] * world,) + local[1:]
    probe = make_config(name, shape=gshape).forward_ast_gpu
    halo = max(probe.halo(probe.input_fields[0].name)[0])
    dtype = probe.input_fields[0].dtype.numpy_dtype
    results = {}
    for fuse in (False, True):
        dh = SlabDataHandling(gshape, rank, world, 2 * halo, dev, peer_halo=peer and world > 1)
        dh.add_arrays('u, out', dtype=dtype)
        kernel = CompiledKernel(make_config(name, shape=dh.dec.local_shape).forward_ast_gpu)
        gen = torch.Generator(device=dev)
        gen.manual_seed(100 + rank)
        dh.owned('u').copy_(torch.randn(local, generator=gen, device=dev, dtype=dh.gpu_arrays['u'].dtype))
        start_state = dh.owned('u').clone()
        dh.run_steps(kernel, steps, fuse=fuse)               # warm-up (NVRTC, module load, NCCL connections)
        dh.run_steps(kernel, steps, fuse=fuse)
        dh.owned('u').copy_(start_state)
        if dh.peer is not None:
            dh.peer.dirty = True
        del start_state
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        t0.record()
        for _ in range(reps):
            dh.run_steps(kernel, steps, fuse=fuse)
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1) / (reps * steps)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        results[fuse] = (float(ms), dh.owned('u').clone() if steps * reps <= 64 else None)
        peer_errors = dh.peer.errors() if dh.peer is not None else None
        dh.close()
        del dh
        torch.cuda.empty_cache()
    diff = None
    if results[False][1] is not None:
        a, b = results[False][1], results[True][1]
        diff = float((a - b).abs().max() / a.abs().max().clamp_min(1e-30))
    if rank == 0:
        cells = int(np.prod(gshape))
        for fuse in (False, True):
            ms = results[fuse][0]
            print(json.dumps({'workload': name, 'n_gpus': world, 'per_gpu_shape': list(local), 'steps': steps, 'fused_pairs': fuse,
                              'ms_per_step': ms, 'gcell_steps_per_s': cells / ms / 1e6,
                              'speedup_vs_single': results[False][0] / ms, 'rel_diff_fused_vs_single': diff,
                              'halo': 'peer' if peer and world > 1 else 'nccl', 'peer_wait_timeouts': peer_errors}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
