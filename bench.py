#!/usr/bin/env python
"""bench.py — forward+adjoint stencil throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4|c5]

One "step" = one forward kernel + one adjoint kernel over the whole field (per GPU: weak scaling, every rank owns
a slab of the workload's full single-GPU shape; ghost planes are exchanged with the neighbours before each kernel
when N > 1).  Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for how each number is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'forward+adjoint Mcell-updates/s'
UNIT = 'Mcell-updates/s'
WORKLOADS = {
    'c2': 'C2: 2-D 5-point diffusion step + adjoint, 8192x8192 fp32, zeros boundary',
    'c3': 'C3: 3-D 7-point heat-equation stencil + adjoint, 1024^3 fp32 per GPU (slab-sharded along dim 0), zeros boundary',
    'c4': 'C4: 3-D 27-point stencil + adjoint, 768^3 fp64, zeros boundary',
    'c5': 'C5: 2-D TV-denoising gradient + adjoint, batch 16 of 4096x4096 fp32, zeros boundary',
}
DTYPE = {'c2': 'f32', 'c3': 'f32', 'c4': 'f64', 'c5': 'f32'}
CPU_SAMPLE_SHAPE = {'c2': (4096, 4096), 'c3': (384, 384, 384), 'c4': (192, 192, 192), 'c5': (2, 2048, 2048)}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--shape', type=int, nargs='*', default=None, help='override the per-GPU field shape')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--strong', action='store_true',
                    help='strong scaling: the workload shape is the GLOBAL field, split along dim 0 over the ranks '
                         '(default: weak scaling, every rank owns the full workload shape)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.lines = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.lines.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nme, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(nme)
        sm.sort()
        # the median over the samples under load (the upper half: the sampler also sees the idle edges)
        load = sm[len(sm) // 2:] if sm else []
        return {'sm_mhz': load[len(load) // 2] if load else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_run(workload, steps, warmup, shape=None):
    """The reference's CPU implementation of the path, restated (oracle/cgen.py, 'fast' flavour = pystencils'
    cpujit flag set, OpenMP over all host threads), timed on a bounded sample of the workload."""
    import numpy as np
    from oracle.cgen import compile_c
    from pystencils_autodiff_b200.configs import make_config

    shape = tuple(shape or CPU_SAMPLE_SHAPE[workload])
    threads = host_threads()
    os.environ['OMP_NUM_THREADS'] = str(threads)
    op = make_config(workload, shape=shape, boundary_handling='zeros')
    fwd = compile_c(op.forward_assignments, 'zeros', op.op_name + '_forward_cpu', 'fast')
    bwd = compile_c(op.backward_assignments, 'zeros', op.op_name + '_backward_cpu', 'fast')
    rng = np.random.default_rng(0)
    arrays = {}
    for f in set(op.forward_fields) | set(op.backward_fields):
        arrays[f.name] = rng.uniform(0.1, 1.0, size=shape).astype(f.dtype.numpy_dtype)
    cells = int(np.prod(shape))

    def step():
        fwd(**{n: arrays[n] for n in fwd.field_names})
        bwd(**{n: arrays[n] for n in bwd.field_names})

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    return dict(value=cells / dt / 1e6, unit=UNIT, cores=threads, kind='port',
                sample='%s: forward+adjoint over a %s sample (%d cells), %d timed passes, gcc -Ofast -march=native -fopenmp, '
                       'restated pystencils CPU path (pystencils itself is not installable here)'
                       % (workload, 'x'.join(map(str, shape)), cells, steps)), dt


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    base, dt = cpu_reference_run(args.workload, args.steps, args.warmup, args.shape)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': DTYPE[args.workload], 'data': 'synthetic',
        'config': {'workload': WORKLOADS[args.workload], 'cpu_sample': base['sample']},
        'cpu_baseline': base,
        'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------
def main_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pystencils_autodiff_b200 import runtime
    from pystencils_autodiff_b200.configs import CONFIG_SHAPES, make_config
    from pystencils_autodiff_b200.datahandling import SlabStencilOp

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl ours needs a CUDA device: this backend has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    wl = args.workload
    shape = tuple(args.shape or CONFIG_SHAPES[wl]['shape'])   # per-GPU (local) shape: weak scaling
    if args.strong:
        if shape[0] % world:
            raise SystemExit('--strong: dim 0 (%d) must be divisible by the number of GPUs (%d)' % (shape[0], world))
        shape = (shape[0] // world,) + shape[1:]
    t_setup0 = time.perf_counter()
    op = make_config(wl, shape=shape, boundary_handling='zeros')
    slab = SlabStencilOp(op, local_shape=shape, rank=rank, world_size=world, device=dev)
    setup_ms = {'symbolic_and_emit_ms': (time.perf_counter() - t_setup0) * 1e3}
    cells = int(np.prod(shape))
    b_fwd = op.forward_ast_gpu.bytes_per_cell()
    b_bwd = op.backward_ast_gpu.bytes_per_cell()

    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    slab.randomize(g)

    def step(events=None):
        if events is not None:
            events[0].record()
        slab.forward()
        if events is not None:
            events[1].record()
        slab.backward()
        if events is not None:
            events[2].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the first step pays for NVRTC (or a cubin-cache hit), module load and — N > 1 — the NCCL connections: reported
    # separately, never inside a timed region
    t_first0 = time.perf_counter()
    step()
    barrier()
    setup_ms['first_step_ms'] = (time.perf_counter() - t_first0) * 1e3
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    # per-kernel durations: CUDA events around the forward and the adjoint launch on every `stride`-th timed step (an
    # event between two 90 us kernels costs a few us of GPU idle time, so not on every step)
    stride = 1 if args.steps < 40 else 8

    def timed_region():
        n0 = runtime.launch_count()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
            time.sleep(0.3)
        evs = {i: [torch.cuda.Event(enable_timing=True) for _ in range(3)] for i in range(0, args.steps, stride)}
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        t_host0 = time.perf_counter()
        for i in range(args.steps):
            step(evs.get(i))
        host_ms_ = (time.perf_counter() - t_host0) * 1e3 / args.steps   # CPU time spent issuing one step
        end.record()
        barrier()
        clk_ = clocks.stop() if rank == 0 else None
        return (start.elapsed_time(end), sum(e[0].elapsed_time(e[1]) for e in evs.values()) / len(evs),
                sum(e[1].elapsed_time(e[2]) for e in evs.values()) / len(evs), host_ms_, clk_,
                runtime.launch_count() - n0, len(evs))

    ms_total, t_fwd, t_bwd, host_ms, clk, launches, n_samples = timed_region()
    # a run that saw a hardware / thermal slowdown is rejected and measured again, once (sw_power_cap is kept and noted)
    bad = {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    retry = torch.tensor([1 if (rank == 0 and clk and bad & set(clk.get('reasons', []))) else 0], device=dev)
    if world > 1:
        dist.all_reduce(retry, op=dist.ReduceOp.MAX)
    if int(retry.item()):
        first_reasons = clk.get('reasons') if clk else None
        time.sleep(2.0)
        ms_total, t_fwd, t_bwd, host_ms, clk, launches, n_samples = timed_region()
        if clk is not None:
            clk['remeasured_after'] = first_reasons
    if world > 1:
        t = torch.tensor([ms_total, t_fwd, t_bwd], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, t_fwd, t_bwd = [float(v) for v in t.tolist()]
    ms_step = ms_total / args.steps
    value = cells * world / (ms_step * 1e-3) / 1e6

    # ---- forward and adjoint as ONE launch (AutoDiffOp.fused_kernel_gpu): possible here because the upstream gradient
    # is an input of the step; reported beside the headline, not instead of it
    fused_ms = None
    if world == 1:
        try:
            fk = op.fused_kernel_gpu
            arrs = {f.name: slab.dh.gpu_arrays[f.name] for f in fk.fields}
            for _ in range(3):
                fk(**arrs)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            f0.record()
            nf = max(5, args.steps // 4)
            for _ in range(nf):
                fk(**arrs)
            f1.record()
            torch.cuda.synchronize()
            fused_ms = f0.elapsed_time(f1) / nf
        except NotImplementedError:
            fused_ms = None

    # ---- two unrolled steps of the forward stencil as ONE launch (emit_chain.py) beside two
    # single-step launches; reported next to the headline, not part of it
    steps_info = None
    if world == 1:
        try:
            fk = slab.fwd
            if fk.fused_steps_reason() is None:
                fin, fout = op.forward_ast_gpu.input_fields[0].name, op.forward_ast_gpu.output_fields[0].name
                g = slab.dh.gpu_arrays
                src, dst, tmp = g[fin], g[fout], g[[n for n in g if n not in (fin, fout)][-1]]

                def _time(fn, n=5):
                    for _ in range(3):
                        fn()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    a.record()
                    for _ in range(n):
                        fn()
                    b.record()
                    torch.cuda.synchronize()
                    return a.elapsed_time(b) / n

                def _two():
                    fk(**{fin: src, fout: tmp})
                    fk(**{fin: tmp, fout: dst})
                t_two = _time(_two)
                t_x2 = _time(lambda: fk(**{fin: src, fout: dst}, _variant='march_x2'))
                steps_info = {'two_launches_ms': t_two, 'one_fused_launch_ms': t_x2, 'speedup': t_two / t_x2,
                              'gcell_steps_per_s': 2 * src.numel() / (t_x2 * 1e-3) / 1e9,
                              'used_by_default': bool(src.element_size() == 4),
                              'note': 'out = S(S(u)) with one read and one write of the field; run_steps() / '
                                      'create_unrolled_torch_op() fuse pairs only where this is a win (4-byte fields)'}
        except Exception as exc:   # a diagnostic beside the headline must never take the line down
            steps_info = {'error': '%s: %s' % (type(exc).__name__, exc)}

    # ---- end to end through the public API with HOST buffers (copies inside the timed region) -------------------
    e2e_error = None
    try:
        e2e = slab.end_to_end(args.e2e_steps, barrier)
    except Exception as exc:   # e.g. not enough pinnable host memory: keep the line, say what happened
        e2e_error = '%s: %s' % (type(exc).__name__, exc)
        e2e = {'ms_per_step': float('inf'), 'h2d': 0, 'd2h': 0}
    if world > 1:
        t = torch.tensor([e2e['ms_per_step']], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e['ms_per_step'] = float(t.item())
    if e2e_error is None and e2e.get('matches_resident') is False:
        e2e_error = 'streamed results differ from the resident kernels on planes %s' % e2e.get('checked_planes')
    e2e_value = cells * world / (e2e['ms_per_step'] * 1e-3) / 1e6 if e2e_error is None else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peaks()
    achieved = cells * b_fwd / (t_fwd * 1e-3) / 1e9
    pair = cells * (b_fwd + b_bwd) / ((t_fwd + t_bwd) * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(wl, {}).get('forward_dram_bytes_per_launch')
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong' if args.strong else 'weak', 'vs_baseline': None,
        'dtype': DTYPE[wl], 'data': 'synthetic',
        'config': {'workload': WORKLOADS[wl], 'per_gpu_shape': list(shape), 'cells_per_gpu': cells,
                   'bytes_per_cell': {'forward': b_fwd, 'adjoint': b_bwd},
                   'l2': 'inputs larger than L2 (%.1f GB per field vs 126 MB): no flush needed'
                         % (cells * op.forward_input_fields[0].dtype.itemsize / 1e9),
                   'kernel_variants': slab.variants(), 'halo_exchange': slab.exchange_kind if world > 1 else 'none (1 GPU)'},
        'roofline': {'bound': 'hbm', 'kernel': op.forward_ast_gpu.function_name, 'achieved': achieved, 'peak': peak,
                     'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                     'forward_ms': t_fwd, 'adjoint_ms': t_bwd, 'kernel_timing_samples': n_samples,
                     'pair_achieved': pair, 'pair_frac': pair / peak,
                     'pair_frac_of_8000_nominal': pair / 8000.0},
        'clocks': clk,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': e2e['h2d'], 'd2h_bytes_per_step': e2e['d2h'],
                'ms_per_step': e2e['ms_per_step'] if e2e_error is None else None, 'steps': args.e2e_steps, 'error': e2e_error,
                'matches_resident': e2e.get('matches_resident'), 'checked_planes': e2e.get('checked_planes'),
                'api': ('HostStreamedOp(AutoDiffOp)(host_in, host_out): pinned host fields streamed through the GPU in plane '
                        'chunks, H2D / forward+adjoint kernels / D2H overlapped on three streams') if world == 1 else
                       'SlabStencilOp: H2D of the slab, halo exchange + kernels, D2H of outputs and input gradients; upload, '
                       'compute and download streams (the upstream gradients go up while the outputs come down)'},
        'gpu_launches': launches,
        'setup': setup_ms,
        'host_issue_ms_per_step': host_ms,
        'fused_steps': steps_info,
        'fused_forward_adjoint': None if fused_ms is None else {
            'ms_per_step': fused_ms, 'value': cells / (fused_ms * 1e-3) / 1e6, 'unit': UNIT,
            'bytes_per_cell': op.fused_ast_gpu.bytes_per_cell(),
            'note': 'one launch over the union of forward and adjoint assignments (outside the headline timed region)'},
    }
    if not args.no_cpu_baseline:
        try:
            base, _ = cpu_reference_run(wl, steps=5, warmup=2)
        except Exception as exc:
            base = {'value': None, 'unit': UNIT, 'cores': host_threads(), 'kind': 'port', 'sample': 'failed',
                    'error': '%s: %s' % (type(exc).__name__, exc)}
        line['cpu_baseline'] = base
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == 'reference' else main_ours(a))
