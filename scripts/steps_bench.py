"""Two fused steps per launch (emit_chain.py) against two single-step launches: agreement on a small grid and at full
size, and timing over a few tile geometries.
    python scripts/steps_bench.py c4            # 27-point fp64, 768^3
    python scripts/steps_bench.py c3 "2,30,4,0;2,22,4,0"   # candidates as ry,ty,sx,lookahead[,exchange[,min_ctas]]
    python scripts/steps_bench.py c2                        # 2-D 'zeros' stencils: lifted to one-plane 3-D fields
    python scripts/steps_bench.py c4 "3,21,4,0,1" c3 "2,30,4,0,1"      # several workloads in one process"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pystencils_autodiff_b200.configs import make_config, CONFIG_SHAPES
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel, numpy_dtype_to_torch
from pystencils_autodiff_b200.emit import MarchTuning

DEFAULT_CANDIDATES = {'c3': '0,0,0,0;2,22,4,0;4,28,4,0;3,33,4,0', 'c4': '0,0,0,0;4,28,2,0;3,21,2,0;2,22,2,0',
                      'c2': '0,0,0,0;2,30,4,0,1;2,30,4,0,0;4,44,4,0,1'}


def timed(fn, iters=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def _tuning(cand):
    v = [int(x) for x in cand.split(',')]
    return MarchTuning(ry=v[0], ty=v[1], sx=v[2], lookahead=v[3], exchange=bool(v[4]) if len(v) > 4 else None,
                       min_ctas=v[5] if len(v) > 5 else 0)


def main():
    args = sys.argv[1:] or ['c4']
    i = 0
    while i < len(args):
        name = args[i]
        cands = DEFAULT_CANDIDATES[name]
        if i + 1 < len(args) and args[i + 1][0].isdigit():
            cands = args[i + 1]
            i += 1
        i += 1
        run(name, cands.split(';'))


def run(name, cands):
    dev = torch.device('cuda:0')
    # ---- small grid, both boundary modes, forward and adjoint kernels: one fused launch against two single-step launches
    # (parity against the CPU oracle is the tests' job: tests/test_gpu_steps.py, tests/test_march_replay.py)
    nd = len(CONFIG_SHAPES[name]['shape'])
    small = (19, 45, 252)[-nd:]
    for bh in ('zeros', None):
        op = make_config(name, shape=small, boundary_handling=bh)
        for ir in (op.forward_ast_gpu, op.backward_ast_gpu):
            k = CompiledKernel(ir)
            k.tuning_x2 = _tuning(cands[0])          # the first candidate's geometry
            if k.fused_steps_reason():
                print('small %-28s bh=%-5s not fusable: %s' % (ir.function_name, bh, k.fused_steps_reason()), flush=True)
                continue
            fin, fout = ir.input_fields[0], ir.output_fields[0]
            ut = torch.randn(small, dtype=numpy_dtype_to_torch(fin.dtype.numpy_dtype), device=dev)
            a, b, out = torch.empty_like(ut), torch.empty_like(ut), torch.full_like(ut, float('nan'))
            k(**{fin.name: ut, fout.name: a})
            k(**{fin.name: a, fout.name: b})
            k(**{fin.name: ut, fout.name: out}, _variant='march_x2')
            err = float((out - b).abs().max())
            err5 = float((k.run_steps(ut, 5, fuse=True) - k.run_steps(ut, 5, fuse=False)).abs().max())
            print('small %-28s bh=%-5s fused pair vs two launches %.3g   5 steps fused vs unfused %.3g'
                  % (ir.function_name, bh, err, err5), flush=True)
    # ---- full size
    shape = CONFIG_SHAPES[name]['shape']
    op = make_config(name, shape=shape)
    ir = op.forward_ast_gpu
    fin, fout = ir.input_fields[0], ir.output_fields[0]
    dt = numpy_dtype_to_torch(fin.dtype.numpy_dtype)
    cells = int(np.prod(shape))
    u = torch.randn(shape, dtype=dt, device=dev)
    a = torch.empty_like(u)
    b = torch.empty_like(u)
    k1 = CompiledKernel(ir)

    def two_single():
        k1(**{fin.name: u, fout.name: a})
        k1(**{fin.name: a, fout.name: b})
    t1 = timed(two_single)
    bpc = ir.bytes_per_cell()
    print('%s %s: two single-step launches %.3f ms (%.0f GB/s algorithmic per launch)' % (name, shape, t1, 2 * cells * bpc / t1 / 1e6),
          flush=True)
    ref = b.clone()
    for cand in cands:
        k2 = CompiledKernel(ir)
        k2.tuning_x2 = _tuning(cand)
        try:
            ek = k2.emitted('march_x2')
        except ValueError as e:
            print('  %-12s not emitted: %s' % (cand, e))
            continue
        out = torch.empty_like(u)
        t2 = timed(lambda: k2(**{fin.name: u, fout.name: out}, _variant='march_x2'))
        diff = float((out - ref).abs().max())
        attrs = k2.native('march_x2').attributes()
        print('  %s %-12s tile %dx%d ry=%d sx=%d stages=%d regs=%d occ=%d: %.3f ms  = %.2fx two launches, %.1f Gcell-steps/s, '
              '%.0f GB/s of field traffic, max|diff| vs two launches %.3g'
              % ('x2e' if ek.geometry['exchange'] else 'x2 ', cand, ek.geometry['TY'], ek.geometry['TX'], ek.geometry['RY'], ek.geometry['SX'], ek.geometry['STAGES'],
                 attrs['num_regs'], attrs['max_ctas_per_sm'], t2, t1 / t2, 2 * cells / t2 / 1e6, cells * bpc / t2 / 1e6, diff), flush=True)


if __name__ == '__main__':
    main()
