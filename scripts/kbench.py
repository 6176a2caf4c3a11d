"""Quick per-kernel timing (CUDA events on the launching stream) for tuning: python scripts/kbench.py c3 512 512 512"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pystencils_autodiff_b200.configs import make_config, CONFIG_SHAPES
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel, numpy_dtype_to_torch
from pystencils_autodiff_b200.emit import MarchTuning


def time_kernel(k, tensors, scalars, iters=10, warm=3, variant=None):
    for _ in range(warm):
        k(**tensors, **scalars, _variant=variant)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        k(**tensors, **scalars, _variant=variant)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    name = sys.argv[1]
    shape = tuple(int(s) for s in sys.argv[2:]) or CONFIG_SHAPES[name]['shape']
    tun = {}
    for kv in os.environ.get('PSAD_TUNE', '').split(','):
        if kv:
            k, v = kv.split('=')
            tun[k] = bool(int(v)) if k in ("carry", "shuffle", "plane_sums", "arrival", "linopt", "cross_cse", "lds_pair", "exchange") else int(v)
    tuning = MarchTuning(**tun) if tun else None
    extra = {'fast_math': True} if os.environ.get('PSAD_FAST_MATH') else {}
    if os.environ.get('PSAD_ADJOINT_MODE'):          # 'exact': the true transpose for non-linear stencils (C5)
        extra['adjoint_mode'] = os.environ['PSAD_ADJOINT_MODE']
    if os.environ.get('PSAD_BH'):
        extra['boundary_handling'] = None if os.environ['PSAD_BH'] == 'none' else os.environ['PSAD_BH']
    op = make_config(name, shape=shape, **extra)
    dev = torch.device('cuda:0')
    cells = 1
    for s in shape:
        cells *= s
    for ir in (op.forward_ast_gpu, op.backward_ast_gpu):
        k = CompiledKernel(ir, tuning)
        tensors = {f.name: torch.rand(shape, dtype=numpy_dtype_to_torch(f.dtype.numpy_dtype), device=dev) + 0.5
                   for f in k.fields}
        scal = {s: 1.0 for s in k.scalars}
        for variant in (["march"] if os.environ.get("PSAD_MARCH_ONLY") and "march" in k.variants else k.variants):
            med, best = time_kernel(k, tensors, scal, variant=variant)
            bpc = ir.bytes_per_cell()
            attrs = k.native(variant).attributes()
            print('%-26s %-8s %s  median %.3f ms best %.3f ms  %.1f Gcell/s  %.0f GB/s (algorithmic %d B/cell)  regs=%d occ=%d'
                  % (ir.function_name, variant, shape, med, best, cells / med / 1e6, cells * bpc / med / 1e6, bpc,
                     attrs['num_regs'], attrs['max_ctas_per_sm']), flush=True)


if __name__ == '__main__':
    main()
