"""The C-ABI shared library: loads without a GPU, exports every symbol include/psad.h declares, compiles cubins for
sm_100a with NVRTC on a CPU-only machine, and fails loudly (no fallback) when asked to run without a driver."""
import ctypes
import os
import re

import pytest

from pystencils_autodiff_b200 import runtime
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    with open(os.path.join(ROOT, 'include', 'psad.h')) as fh:
        text = fh.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(psad_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(runtime.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), 'libpsad.so does not export %s' % n


def test_abi_version_and_struct_layout():
    L = runtime.lib()
    assert L.psad_abi_version() == runtime.PSAD_ABI_VERSION
    # psad_field_plan_t: 12 int32; psad_plan_t: 21 int32 + 12 field plans; psad_field_arg_t: ptr + 7 int64
    assert ctypes.sizeof(runtime.FieldPlan) == 12 * 4
    assert ctypes.sizeof(runtime.Plan) == 21 * 4 + 12 * 48
    assert ctypes.sizeof(runtime.FieldArg) == 8 + 7 * 8
    assert ctypes.sizeof(runtime.Range) == 12 * 8


def test_nvrtc_compiles_sm100a_cubin_without_gpu(tmp_path):
    op = make_config('c3', shape=(16, 16, 128))
    k = CompiledKernel(op.forward_ast_gpu)
    assert k.variants == ['generic', 'march']
    for v in k.variants:
        ek = k.emitted(v)
        key = ek.cache_key + '_t'
        hit, log = runtime.compile_source(ek.source, key, list(ek.options) + ['--ptxas-options=-v'])
        path = runtime.cubin_path(key)
        assert os.path.getsize(path) > 1000
        with open(path, 'rb') as fh:
            assert fh.read(4) == b'\x7fELF'
        if not hit:
            assert 'sm_100a' in log
        hit2, _ = runtime.compile_source(ek.source, key, list(ek.options) + ['--ptxas-options=-v'])
        assert hit2
        os.remove(path)


def test_compile_error_is_reported_not_swallowed():
    with pytest.raises(RuntimeError) as e:
        runtime.compile_source('extern "C" __global__ void k() { this is not cuda; }', 'broken_kernel_test')
    assert 'NVRTC' in str(e.value) and 'error' in str(e.value)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    op = make_config('c2', shape=(16, 32))
    k = CompiledKernel(op.forward_ast_gpu)
    with pytest.raises(RuntimeError) as e:
        k.native('generic')
    assert 'libcuda' in str(e.value) or 'driver' in str(e.value).lower()
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    with pytest.raises(Exception):
        fn.apply(torch.zeros(16, 32))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'pystencils_autodiff_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cpp', '.cuh', '.h')):
                with open(os.path.join(dirpath, f)) as fh:
                    text = fh.read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f


def test_plan_launch_work_decomposition():
    """psad_plan_launch (the device-free half of psad_kernel_launch): parameter block and persistent-grid work list of
    march kernels — tiles cover the array, chunks cover the written planes, grid = min(items, SMs x resident CTAs)."""
    import ctypes
    import struct
    import numpy as np
    from pystencils_autodiff_b200 import configs
    from pystencils_autodiff_b200.emit import emit_march
    from pystencils_autodiff_b200.emit_chain import emit_march_chain
    L = runtime.lib()
    shape = (300, 70, 1000)
    op = configs.heat3d_op(shape=shape)
    for ek, halo in ((emit_march(op.forward_ast_gpu), 1), (emit_march_chain(op.forward_ast_gpu), 2)):
        plan = runtime.make_plan(ek.plan)
        fa = (runtime.FieldArg * 2)()
        buf = np.zeros(64, dtype=np.uint8)
        base = buf.ctypes.data + (-buf.ctypes.data) % 16
        for i in range(2):
            fa[i].ptr = base                       # never dereferenced: planning only
            fa[i].shape[:] = shape
            fa[i].stride[:] = [shape[1] * shape[2], shape[2], 1, 0]
        L.psad_args_size.restype = ctypes.c_size_t
        nbytes = int(L.psad_args_size())
        assert nbytes >= 752 and nbytes % 8 == 0
        args = ctypes.create_string_buffer(nbytes)
        grid = (ctypes.c_uint * 3)()
        for sms, ctas in ((148, 1), (4, 2)):
            rc = L.psad_plan_launch(ctypes.byref(plan), sms, ctas, fa, 2, None, 0, None, args, nbytes, grid)
            assert rc == 0, L.psad_last_error()
            n_items, = struct.unpack_from('q', args.raw, 728)
            tiles_x, tiles_y, n_chunks, chunk = struct.unpack_from('4i', args.raw, 736)
            assert tiles_x == -(-shape[2] // ek.plan['tile_x']) and tiles_y == -(-shape[1] // ek.plan['tile_y'])
            assert n_items == tiles_x * tiles_y * n_chunks
            assert chunk * n_chunks >= shape[0] > chunk * (n_chunks - 1)
            assert grid[0] == min(n_items, sms * ctas) and grid[1] == grid[2] == 1
            # the busiest CTA's steps stay within 15 % of a perfectly balanced split (chunks amortise warm-up planes)
            steps = -(-n_items // (sms * ctas)) * (chunk + 2 * halo)
            assert steps <= 1.15 * (tiles_x * tiles_y * shape[0] / (sms * ctas)) + chunk + 2 * halo
        # wrong field count and wrong block size are rejected with a message
        assert L.psad_plan_launch(ctypes.byref(plan), 148, 1, fa, 1, None, 0, None, args, nbytes, grid) != 0
        assert b'expected 2 fields' in L.psad_last_error()
        assert L.psad_plan_launch(ctypes.byref(plan), 148, 1, fa, 2, None, 0, None, args, 100, grid) != 0
        # launch ranges (slabs): the chunks cover the WRITTEN planes only — also for fused-step kernels — and a write
        # range outside the array is rejected
        rng = runtime.Range()
        for d in range(3):
            rng.iter_hi[d] = rng.write_hi[d] = shape[d]
        rng.write_lo[0], rng.write_hi[0] = 2 * halo, 2 * halo + 40
        rc = L.psad_plan_launch(ctypes.byref(plan), 148, 1, fa, 2, None, 0, ctypes.byref(rng), args, nbytes, grid)
        assert rc == 0, L.psad_last_error()
        tiles_x, tiles_y, n_chunks, chunk = struct.unpack_from('4i', args.raw, 736)
        assert chunk * n_chunks >= 40 > chunk * (n_chunks - 1)
        rng.write_hi[0] = shape[0] + 1
        assert L.psad_plan_launch(ctypes.byref(plan), 148, 1, fa, 2, None, 0, ctypes.byref(rng), args, nbytes, grid) != 0
        assert b'outside the array extent' in L.psad_last_error()
