"""ctypes binding of the C ABI in ``include/psad.h`` (``csrc/libpsad.so``).

This is the only place Python talks to native code.  There is no fallback: if the shared library has not been
built, or a call fails, a ``RuntimeError`` carrying ``psad_last_error()`` is raised.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, 'csrc')
KERNEL_DIR = os.path.join(CSRC_DIR, 'kernels')
LIB_PATH = os.path.join(CSRC_DIR, 'libpsad.so')


def _default_cache_dir():
    """$PSAD_CACHE_DIR, else the in-tree cache (it travels with the repo snapshot), else ~/.cache when the package
    directory is read-only (site-packages installs)."""
    env = os.environ.get('PSAD_CACHE_DIR')
    if env:
        return env
    in_tree = os.path.join(_HERE, '_cubin_cache')
    try:
        os.makedirs(in_tree, exist_ok=True)
        if os.access(in_tree, os.W_OK):
            return in_tree
    except OSError:
        pass
    return os.path.join(os.path.expanduser('~'), '.cache', 'pystencils_autodiff_b200', 'cubin')


DEFAULT_CACHE_DIR = _default_cache_dir()

PSAD_MAX_FIELDS = 12
PSAD_MAX_SCALARS = 16
PSAD_ABI_VERSION = 1


class FieldPlan(ctypes.Structure):
    _fields_ = [('elem_size', ctypes.c_int32), ('is_input', ctypes.c_int32), ('is_output', ctypes.c_int32),
                ('index_size', ctypes.c_int32), ('tma', ctypes.c_int32), ('box', ctypes.c_int32 * 3),
                ('reserved', ctypes.c_int32 * 4)]


class Plan(ctypes.Structure):
    _fields_ = [('abi_version', ctypes.c_int32), ('kind', ctypes.c_int32), ('ndim', ctypes.c_int32),
                ('n_fields', ctypes.c_int32), ('n_scalars', ctypes.c_int32), ('threads', ctypes.c_int32),
                ('smem_bytes', ctypes.c_int32), ('tile_x', ctypes.c_int32), ('tile_y', ctypes.c_int32),
                ('chunk', ctypes.c_int32), ('ctas_per_sm', ctypes.c_int32), ('boundary', ctypes.c_int32),
                ('ghost_layers', ctypes.c_int32), ('reserved', ctypes.c_int32 * 8),
                ('field', FieldPlan * PSAD_MAX_FIELDS)]


class FieldArg(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p), ('shape', ctypes.c_int64 * 3), ('stride', ctypes.c_int64 * 4)]


class Peer(ctypes.Structure):
    """``psad_peer_t``: the neighbouring GPUs' arrays and completion counters of a peer-halo launch."""
    _fields_ = [('lo_ptr', ctypes.c_void_p * PSAD_MAX_FIELDS), ('hi_ptr', ctypes.c_void_p * PSAD_MAX_FIELDS),
                ('lo_planes', ctypes.c_int64), ('hi_planes', ctypes.c_int64), ('flag_lo', ctypes.c_void_p),
                ('flag_hi', ctypes.c_void_p), ('error_flag', ctypes.c_void_p), ('expect', ctypes.c_uint32),
                ('ghost_planes', ctypes.c_int32)]


class Range(ctypes.Structure):
    _fields_ = [('iter_lo', ctypes.c_int64 * 3), ('iter_hi', ctypes.c_int64 * 3),
                ('write_lo', ctypes.c_int64 * 3), ('write_hi', ctypes.c_int64 * 3)]


_lib = None
_lock = threading.Lock()


def build_library(force=False, verbose=False):
    """Compile ``csrc/psad_runtime.cpp`` into ``csrc/libpsad.so`` (plain g++: all device code goes through NVRTC)."""
    src = os.path.join(CSRC_DIR, 'psad_runtime.cpp')
    deps = [src, os.path.join(_HERE, '..', 'include', 'psad.h'), os.path.join(KERNEL_DIR, 'psad_args.h')]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = ['g++', '-O2', '-fPIC', '-shared', '-std=c++17', '-Wall', '-o', LIB_PATH + '.tmp', src, '-ldl', '-lpthread']
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    if os.path.exists(os.path.join(cuda_inc, 'nvtx3', 'nvToolsExt.h')):     # header-only NVTX ranges (optional)
        cmd += ['-isystem', cuda_inc]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    os.replace(LIB_PATH + '.tmp', LIB_PATH)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:             # the common case takes no lock
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'pystencils_autodiff_b200: native runtime %s is missing. Build it with '
                '`python -c "import __graft_entry__ as g; g.build()"` (there is no CPU/PyTorch fallback).' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.psad_last_error.restype = ctypes.c_char_p
        L.psad_launch_count.restype = ctypes.c_uint64
        L.psad_init.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.psad_compile.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int,
                                   ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_void_p)]
        L.psad_free.argtypes = [ctypes.c_void_p]
        L.psad_kernel_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p,
                                         ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.POINTER(Plan),
                                         ctypes.POINTER(ctypes.c_void_p)]
        L.psad_kernel_destroy.argtypes = [ctypes.c_void_p]
        L.psad_kernel_attributes.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int)] * 4
        L.psad_kernel_launch.argtypes = [ctypes.c_void_p, ctypes.POINTER(FieldArg), ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.POINTER(Range),
                                         ctypes.c_void_p]
        L.psad_plan_launch.argtypes = [ctypes.POINTER(Plan), ctypes.c_int, ctypes.c_int, ctypes.POINTER(FieldArg),
                                       ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.POINTER(Range),
                                       ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint)]
        L.psad_kernel_launch_peer.argtypes = [ctypes.c_void_p, ctypes.POINTER(FieldArg), ctypes.c_int,
                                              ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.POINTER(Range),
                                              ctypes.POINTER(Peer), ctypes.c_void_p]
        L.psad_ipc_export.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.psad_ipc_open.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p)]
        L.psad_ipc_close.argtypes = [ctypes.c_void_p]
        L.psad_stream_write_u32.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
        L.psad_peer_wait.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
        L.psad_launch_cache_stats.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)] * 2
        L.psad_device_info.argtypes = [ctypes.POINTER(ctypes.c_int)] * 4 + [ctypes.POINTER(ctypes.c_size_t)]
        L.psad_nccl_unique_id.argtypes = [ctypes.c_void_p]
        L.psad_nccl_comm_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.psad_nccl_comm_destroy.argtypes = [ctypes.c_void_p]
        L.psad_halo_exchange.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        if L.psad_abi_version() != PSAD_ABI_VERSION:
            raise RuntimeError('libpsad.so ABI version mismatch: rebuild it')
        rc = L.psad_init(KERNEL_DIR.encode(), DEFAULT_CACHE_DIR.encode())
        if rc:
            raise RuntimeError('psad_init failed: %s' % L.psad_last_error().decode())
        _lib = L
        return _lib


def check(rc, what=''):
    if rc:
        raise RuntimeError('%s failed (code %d): %s' % (what or 'libpsad call', rc, lib().psad_last_error().decode()))


def make_plan(p):
    plan = Plan()
    plan.abi_version = PSAD_ABI_VERSION
    for k in ('kind', 'ndim', 'n_fields', 'n_scalars', 'threads', 'smem_bytes', 'tile_x', 'tile_y', 'chunk',
              'ctas_per_sm', 'boundary', 'ghost_layers'):
        setattr(plan, k, int(p[k]))
    plan.reserved[0] = int(p.get('warmup', 0))
    plan.reserved[1] = int(p.get('fused_steps', 1))
    plan.reserved[2] = int(p.get('peer', 0))
    for i, f in enumerate(p['fields']):
        fp = plan.field[i]
        for k in ('elem_size', 'is_input', 'is_output', 'index_size', 'tma'):
            setattr(fp, k, int(f[k]))
        for d in range(3):
            fp.box[d] = int(f['box'][d])
    return plan


def _options(options):
    arr = (ctypes.c_char_p * max(1, len(options)))(*[o.encode() for o in options])
    return arr, len(options)


def compile_source(source, cache_key, options=()):
    """NVRTC -> sm_100a cubin in the cache (works without a GPU).  Returns ``(cache_hit, log)``."""
    L = lib()
    arr, n = _options(list(options))
    hit = ctypes.c_int(0)
    log = ctypes.c_void_p()
    rc = L.psad_compile(source.encode(), cache_key.encode(), arr, n, ctypes.byref(hit), ctypes.byref(log))
    text = ''
    if log.value:
        text = ctypes.string_at(log.value).decode(errors='replace')
        L.psad_free(log)
    check(rc, 'psad_compile')
    return bool(hit.value), text


def cubin_path(cache_key):
    return os.path.join(DEFAULT_CACHE_DIR, cache_key + '.cubin')


class NativeKernel:
    """Owns one ``psad_kernel_t``."""

    def __init__(self, emitted):
        L = lib()
        self.emitted = emitted
        self._plan = make_plan(emitted.plan)
        arr, n = _options(list(emitted.options))
        handle = ctypes.c_void_p()
        check(L.psad_kernel_create(emitted.source.encode(), emitted.name.encode(), emitted.cache_key.encode(), arr, n,
                                   ctypes.byref(self._plan), ctypes.byref(handle)), 'psad_kernel_create(%s)' % emitted.name)
        self._handle = handle
        self._scal = (ctypes.c_double * PSAD_MAX_SCALARS)()

    def attributes(self):
        vals = [ctypes.c_int(0) for _ in range(4)]
        check(lib().psad_kernel_attributes(self._handle, *[ctypes.byref(v) for v in vals]), 'psad_kernel_attributes')
        return dict(num_regs=vals[0].value, static_smem=vals[1].value, local_bytes=vals[2].value,
                    max_ctas_per_sm=vals[3].value)

    @staticmethod
    def pack_fields(field_args):
        """``(ptr, shape, strides)`` triples in plan order -> the ``psad_field_arg_t`` array of a launch."""
        n = len(field_args)
        fa = (FieldArg * n)()
        for i, (ptr, shape, strides) in enumerate(field_args):
            fa[i].ptr = ptr
            for d in range(3):
                fa[i].shape[d] = shape[d] if d < len(shape) else 1
            for d in range(4):
                fa[i].stride[d] = strides[d] if d < len(strides) else 0
        return fa

    def launch_packed(self, fa, n, scalars, stream, range_ref, peer=None):
        """A launch whose field array (and range) were packed before: the repeated-launch path.  ``peer``: a ``Peer``
        struct for peer-halo kernels (its ``expect`` set by the caller)."""
        ns = len(scalars)
        scal = (ctypes.c_double * ns)(*scalars) if ns else self._scal
        if peer is None:
            rc = _lib.psad_kernel_launch(self._handle, fa, n, scal, ns, range_ref, stream)
        else:
            rc = _lib.psad_kernel_launch_peer(self._handle, fa, n, scal, ns, range_ref, ctypes.byref(peer), stream)
        if rc:
            check(rc, 'psad_kernel_launch(%s)' % self.emitted.name)

    @staticmethod
    def launch_range_struct(rng):
        """The ``psad_range_t`` of a range dict, built once and kept in the dict (launch ranges are reused every step)."""
        r = rng.get('_ctypes')
        if r is None:
            r = Range()
            for d in range(len(rng['iter_lo'])):
                r.iter_lo[d], r.iter_hi[d] = rng['iter_lo'][d], rng['iter_hi'][d]
                r.write_lo[d], r.write_hi[d] = rng['write_lo'][d], rng['write_hi'][d]
            rng['_ctypes'] = r
        return r

    def launch(self, field_args, scalars, stream, rng=None):
        """field_args: list of ``(ptr, shape, strides)`` in plan order; scalars: list of floats; stream: int handle."""
        n = len(field_args)
        fa = self.pack_fields(field_args)
        # per-call buffer: launches of one kernel from several threads (autograd workers) must not share it
        scal = (ctypes.c_double * max(1, len(scalars)))(*[float(s) for s in scalars]) if scalars else self._scal
        r = None
        if rng is not None:
            r = rng.get('_ctypes') if isinstance(rng, dict) else None
            if r is None:
                r = Range()
                for d in range(len(rng['iter_lo'])):
                    r.iter_lo[d], r.iter_hi[d] = rng['iter_lo'][d], rng['iter_hi'][d]
                    r.write_lo[d], r.write_hi[d] = rng['write_lo'][d], rng['write_hi'][d]
                rng['_ctypes'] = r          # launch ranges are reused every step: build the struct once
            r = ctypes.byref(r)
        check(lib().psad_kernel_launch(self._handle, fa, n, scal, len(scalars), r, ctypes.c_void_p(stream)),
              'psad_kernel_launch(%s)' % self.emitted.name)

    def __del__(self):
        try:
            if self._handle:
                lib().psad_kernel_destroy(self._handle)
        except Exception:
            pass


def ipc_export(ptr):
    """``(handle bytes, offset)`` of a device pointer for another process of this node (``psad_ipc_export``)."""
    handle = (ctypes.c_ubyte * 64)()
    off = ctypes.c_uint64(0)
    check(lib().psad_ipc_export(ctypes.c_void_p(ptr), handle, ctypes.byref(off)), 'psad_ipc_export')
    return bytes(handle), int(off.value)


def ipc_open(handle, offset):
    """Device pointer (int) in THIS process of memory another process exported with :func:`ipc_export`."""
    buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
    out = ctypes.c_void_p()
    check(lib().psad_ipc_open(buf, ctypes.c_uint64(offset), ctypes.byref(out)), 'psad_ipc_open')
    return int(out.value)


def ipc_close(ptr):
    check(lib().psad_ipc_close(ctypes.c_void_p(ptr)), 'psad_ipc_close')


def stream_write_u32(ptr, value, stream):
    rc = lib().psad_stream_write_u32(ctypes.c_void_p(ptr), ctypes.c_uint32(value & 0xffffffff), ctypes.c_void_p(stream))
    if rc:
        check(rc, 'psad_stream_write_u32')


def peer_wait(flag_lo, flag_hi, expect, error_flag, stream):
    """Everything behind this call on ``stream`` starts once the neighbours' launch counters have reached ``expect``."""
    rc = lib().psad_peer_wait(ctypes.c_void_p(flag_lo), ctypes.c_void_p(flag_hi), ctypes.c_uint32(expect & 0xffffffff),
                              ctypes.c_void_p(error_flag), ctypes.c_void_p(stream))
    if rc:
        check(rc, 'psad_peer_wait')


def launch_count():
    return int(lib().psad_launch_count())


def launch_cache_stats():
    """(hits, misses) of the per-kernel launch cache (parameter block + encoded tensor maps) of ``psad_kernel_launch``."""
    h, m = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    lib().psad_launch_cache_stats(ctypes.byref(h), ctypes.byref(m))
    return int(h.value), int(m.value)


def device_info():
    vals = [ctypes.c_int(0) for _ in range(4)]
    smem = ctypes.c_size_t(0)
    check(lib().psad_device_info(*[ctypes.byref(v) for v in vals], ctypes.byref(smem)), 'psad_device_info')
    return dict(device=vals[0].value, sm_count=vals[1].value, cc=(vals[2].value, vals[3].value), smem_optin=smem.value)
