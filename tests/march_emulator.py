"""Replay of an emitted march kernel on the CPU.  TEST INFRASTRUCTURE.

The emitted CUDA source is compiled by g++ against ``tests/cpu_shim/`` (host stand-ins for ``psad_common.cuh`` and
``psad_march.cuh``): the stencil's own ``psad_item_begin`` / ``psad_step`` code — register window, slot rotation,
shuffled halos, masks, vector stores — runs unchanged, one OS thread per lane, over work items decoded by the
product's ``psad_item.cuh`` from a parameter block built by the product's ``psad_plan_launch``.  What is *not*
exercised: TMA, mbarriers and the producer/consumer overlap (GPU tests cover those)."""
import ctypes
import hashlib
import os
import subprocess

import numpy as np

from pystencils_autodiff_b200 import runtime

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, 'cpu_shim')
_SHIM_FULL = os.path.join(_HERE, 'cpu_shim_full')
_SHIM_GENERIC = os.path.join(_HERE, 'cpu_shim_generic')
_BUILD = os.path.join(_HERE, '..', 'oracle', '_build')


class _EmuField(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p), ('stride', ctypes.c_longlong * 3), ('esize', ctypes.c_int),
                ('boxw', ctypes.c_int), ('boxh', ctypes.c_int)]


class _EmuFieldFull(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p), ('stride', ctypes.c_longlong * 3), ('shape', ctypes.c_longlong * 3),
                ('esize', ctypes.c_int), ('boxw', ctypes.c_int), ('boxh', ctypes.c_int)]


def _compile(emitted, full=False):
    os.makedirs(_BUILD, exist_ok=True)
    shim = _SHIM_FULL if full else _SHIM
    h = hashlib.md5(emitted.source.encode())
    for fn in sorted(os.listdir(shim)) + ['psad_item.cuh', 'psad_args.h'] + (['psad_march.cuh'] if full else []):
        path = os.path.join(shim, fn) if os.path.exists(os.path.join(shim, fn)) else os.path.join(runtime.KERNEL_DIR, fn)
        with open(path, 'rb') as fh:
            h.update(fh.read())
    base = os.path.join(_BUILD, 'emu%s_%s_%s' % ('full' if full else '', emitted.name[:32], h.hexdigest()[:12]))
    if not os.path.exists(base + '.so'):
        with open(base + '.%d.cpp' % os.getpid(), 'w') as fh:
            fh.write(emitted.source)
            if full:   # the real psad_march.cuh was included by the source; add the emulated machine + launch loop
                with open(os.path.join(_SHIM_FULL, 'driver.inc')) as inc:
                    fh.write('\n' + inc.read())
        # -ffp-contract=off mirrors -fmad=false; fma()/fmaf() map to the hardware FMA like on the GPU
        subprocess.check_call(['g++', '-std=c++20', '-O1', '-ffp-contract=off', '-mfma', '-fPIC', '-shared', '-pthread', '-w', '-Wno-psabi',
                               '-I', shim, '-I', runtime.KERNEL_DIR, '-o', base + '.so.tmp%d' % os.getpid(), base + '.%d.cpp' % os.getpid()])
        os.replace(base + '.so.tmp%d' % os.getpid(), base + '.so')
    return ctypes.CDLL(base + '.so')


def run(emitted, arrays, scalars=(), sm_count=3, ctas_per_sm=1, launch_range=None, full=False, peer=None):
    """``arrays``: numpy arrays in the plan's field order (outputs first; written in place).

    ``full=False``: the per-step bodies under a host copy of the consumer loop (``cpu_shim/``, warps one after the other
    or — exchange kernels — concurrently).  ``full=True``: the REAL ``psad_march.cuh`` (producer warp, TMA ring, full /
    empty mbarriers) with emulated mbarriers and TMA, every thread of the CTA an OS thread (``cpu_shim_full/``); returns
    ``(ctas, mbarrier waits, TMA loads)``.

    ``peer`` (peer-halo kernels, ``full=True`` only): ``dict(lo=[arrays | None], hi=[arrays | None], ghost_planes=g,
    flags=uint32 array [lower counter, upper counter, error], expect=k)`` — the neighbouring slabs' arrays in plan order
    (None: no neighbour on that side); the parameter block comes from the product's ``psad_plan_launch_peer``."""
    L = runtime.lib()
    plan = runtime.make_plan(emitted.plan)
    n = len(arrays)
    fa = (runtime.FieldArg * n)()
    for i, a in enumerate(arrays):
        assert a.flags['C_CONTIGUOUS'] and a.ctypes.data % 16 == 0
        fa[i].ptr = a.ctypes.data
        for d in range(3):
            fa[i].shape[d] = a.shape[d] if d < a.ndim else 1
            fa[i].stride[d] = a.strides[d] // a.itemsize if d < a.ndim else 0
        fa[i].stride[3] = 0
    sc = (ctypes.c_double * max(1, len(scalars)))(*scalars)
    args = ctypes.create_string_buffer(4096)
    grid = (ctypes.c_uint * 3)()
    rng = None
    if launch_range is not None:
        rng = runtime.Range()
        for d in range(arrays[0].ndim):      # the product's range dicts (datahandling.slab_ranges)
            rng.iter_lo[d], rng.iter_hi[d] = launch_range['iter_lo'][d], launch_range['iter_hi'][d]
            rng.write_lo[d], rng.write_hi[d] = launch_range['write_lo'][d], launch_range['write_hi'][d]
    size = _args_size()
    tma = [(i, f) for i, f in enumerate(emitted.plan['fields']) if f['tma']]
    sets = [arrays]
    if peer is None:
        runtime.check(L.psad_plan_launch(ctypes.byref(plan), sm_count, ctas_per_sm, fa, n, sc, len(scalars),
                                         ctypes.byref(rng) if rng is not None else None, args, size, grid), 'psad_plan_launch')
    else:
        assert full, 'peer kernels are replayed through the real psad_march.cuh only'
        P = runtime.Peer()
        flags = peer['flags']
        for side, nb in (('lo', peer.get('lo')), ('hi', peer.get('hi'))):
            if nb is None:
                continue
            for i, a in enumerate(nb):
                assert a.flags['C_CONTIGUOUS'] and a.shape[1:] == arrays[i].shape[1:]
                getattr(P, side + '_ptr')[i] = a.ctypes.data
            setattr(P, side + '_planes', nb[0].shape[0])
            setattr(P, 'flag_' + side, flags.ctypes.data + (0 if side == 'lo' else 4))
        P.error_flag = flags.ctypes.data + 8
        P.expect = peer['expect']
        P.ghost_planes = peer['ghost_planes']
        L.psad_plan_launch_peer.argtypes = [ctypes.POINTER(runtime.Plan), ctypes.c_int, ctypes.c_int,
                                            ctypes.POINTER(runtime.FieldArg), ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                            ctypes.c_int, ctypes.POINTER(runtime.Range), ctypes.POINTER(runtime.Peer),
                                            ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint)]
        runtime.check(L.psad_plan_launch_peer(ctypes.byref(plan), sm_count, ctas_per_sm, fa, n, sc, len(scalars),
                                              ctypes.byref(rng) if rng is not None else None, ctypes.byref(P), args, size, grid),
                      'psad_plan_launch_peer')
        # a missing neighbour gets the local arrays (never read: the kernel takes that branch only with a counter)
        sets = [arrays, peer.get('lo') or arrays, peer.get('hi') or arrays]
    if grid[0] == 0:
        return 0
    tf = ((_EmuFieldFull if full else _EmuField) * (len(tma) * len(sets)))()
    nd = arrays[0].ndim
    for k, (i, f) in enumerate((i, f) for arrs in sets for (i, f) in tma):
        a = sets[k // len(tma)][i]
        st = [0] * (3 - nd) + [s // a.itemsize for s in a.strides]
        tf[k].ptr = a.ctypes.data
        tf[k].stride[:] = st
        tf[k].esize = a.itemsize
        tf[k].boxw, tf[k].boxh = f['box'][0], f['box'][1]
        if full:
            tf[k].shape[:] = [1] * (3 - nd) + list(a.shape)
    if full:
        so = _compile(emitted, full=True)
        so.psad_emulate_full.argtypes = [ctypes.c_void_p, ctypes.POINTER(_EmuFieldFull), ctypes.c_int, ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_longlong)]
        stats = (ctypes.c_longlong * 2)()
        rc = so.psad_emulate_full(args, tf, len(tma) * len(sets), int(grid[0]), stats)
        assert rc == 0, 'psad_emulate_full failed'
        return int(grid[0]), int(stats[0]), int(stats[1])
    so = _compile(emitted)
    so.psad_emulate.argtypes = [ctypes.c_void_p, ctypes.POINTER(_EmuField), ctypes.c_int, ctypes.c_int]
    rc = so.psad_emulate(args, tf, len(tma), int(grid[0]))
    assert rc == 0, 'psad_emulate failed'
    return int(grid[0])


def run_generic(emitted, arrays, scalars=(), launch_range=None, sm_count=4):
    """Replay of an emitted GENERIC kernel (grid-stride loops, scalar accesses): the source is compiled unchanged against
    ``tests/cpu_shim_generic/`` and called once per CUDA thread over the grid ``psad_plan_launch`` chose.  ``arrays``:
    numpy arrays in plan order, any strides; a trailing index dimension is allowed."""
    L = runtime.lib()
    plan = runtime.make_plan(emitted.plan)
    nd = emitted.plan['ndim']
    n = len(arrays)
    fa = (runtime.FieldArg * n)()
    for i, a in enumerate(arrays):
        fa[i].ptr = a.ctypes.data
        for d in range(3):
            fa[i].shape[d] = a.shape[d] if d < nd else 1
            fa[i].stride[d] = a.strides[d] // a.itemsize if d < nd else 0
        fa[i].stride[3] = a.strides[nd] // a.itemsize if a.ndim > nd else 0
    sc = (ctypes.c_double * max(1, len(scalars)))(*scalars)
    args = ctypes.create_string_buffer(4096)
    grid = (ctypes.c_uint * 3)()
    rng = None
    if launch_range is not None:
        rng = runtime.Range()
        for d in range(nd):
            rng.iter_lo[d], rng.iter_hi[d] = launch_range['iter_lo'][d], launch_range['iter_hi'][d]
            rng.write_lo[d], rng.write_hi[d] = launch_range['write_lo'][d], launch_range['write_hi'][d]
    runtime.check(L.psad_plan_launch(ctypes.byref(plan), sm_count, 1, fa, n, sc, len(scalars),
                                     ctypes.byref(rng) if rng is not None else None, args, _args_size(), grid), 'psad_plan_launch')
    if grid[0] == 0:
        return (0, 0, 0)
    os.makedirs(_BUILD, exist_ok=True)
    h = hashlib.md5(emitted.source.encode())
    for fn in sorted(os.listdir(_SHIM_GENERIC)) + ['psad_args.h']:
        path = os.path.join(_SHIM_GENERIC, fn) if os.path.exists(os.path.join(_SHIM_GENERIC, fn)) else os.path.join(runtime.KERNEL_DIR, fn)
        with open(path, 'rb') as fh:
            h.update(fh.read())
    base = os.path.join(_BUILD, 'emugen_%s_%s' % (emitted.name[:32], h.hexdigest()[:12]))
    if not os.path.exists(base + '.so'):
        with open(base + '.%d.cpp' % os.getpid(), 'w') as fh:
            fh.write(emitted.source)
            with open(os.path.join(_SHIM_GENERIC, 'driver.inc')) as inc:
                fh.write('\n' + inc.read())
        subprocess.check_call(['g++', '-std=c++17', '-O1', '-ffp-contract=off', '-mfma', '-fPIC', '-shared', '-w',
                               '-DPSAD_EMU_KERNEL=' + emitted.name, '-I', _SHIM_GENERIC, '-I', runtime.KERNEL_DIR,
                               '-o', base + '.so.tmp%d' % os.getpid(), base + '.%d.cpp' % os.getpid()])
        os.replace(base + '.so.tmp%d' % os.getpid(), base + '.so')
    so = ctypes.CDLL(base + '.so')
    so.psad_emulate_generic.argtypes = [ctypes.c_void_p] + [ctypes.c_uint] * 4
    rc = so.psad_emulate_generic(args, grid[0], grid[1], grid[2], emitted.plan['threads'])
    assert rc == 0
    return tuple(int(g) for g in grid)


def _args_size():
    """sizeof(PsadArgs) of the library under test (the shims include the same psad_args.h)."""
    L = runtime.lib()
    L.psad_args_size.restype = ctypes.c_size_t
    return int(L.psad_args_size())


def aligned_empty(shape, dtype, fill=None):
    n = int(np.prod(shape))
    raw = np.empty(n * np.dtype(dtype).itemsize + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    a = raw[off:off + n * np.dtype(dtype).itemsize].view(dtype).reshape(shape)
    if fill is not None:
        a[...] = fill
    return a
