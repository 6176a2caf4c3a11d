"""Peer-halo march kernels (``PSAD_PEER``: ghost planes staged from the neighbouring slabs' arrays) replayed on the CPU
through the REAL psad_march.cuh (tests/cpu_shim_full): three slabs of one field, every slab launched once over all of its
owned planes with the neighbours' arrays as peers, against the unsharded oracle.  Local ghost planes next to a neighbour
hold NaN — a load that should have gone to the neighbour but went to the local array poisons the result.  The cross-GPU
ordering itself (counters over NVLink) is a GPU matter: tests/test_gpu_multi.py, scripts/check_peer_halo.py."""
import numpy as np
import pytest

import march_emulator as emu
from oracle.evaluate import evaluate
from pystencils_autodiff_b200 import configs
from pystencils_autodiff_b200.datahandling import slab_ranges
from pystencils_autodiff_b200.emit import MarchTuning, emit_march
from pystencils_autodiff_b200.emit_chain import emit_march_chain


def _slabs(u, sizes, g):
    """Per rank: (start, local array with NaN ghost planes next to a neighbour and zero ghost planes at the global border)."""
    out, start = [], 0
    for r, n in enumerate(sizes):
        a = emu.aligned_empty((n + 2 * g,) + u.shape[1:], u.dtype, np.nan)
        a[g:g + n] = u[start:start + n]
        if r == 0:
            a[:g] = 0
        if r == len(sizes) - 1:
            a[g + n:] = 0
        out.append((start, a))
        start += n
    return out


@pytest.mark.timeout(900)
@pytest.mark.parametrize('make, bh, which, steps, g, sizes, masked, tuning, tol', [
    (configs.heat3d_op, 'zeros', 'forward', 1, 1, (4, 3, 5), False, None, 4e-7),        # the bench's kernel (mask-free)
    (configs.heat3d_op, None, 'backward', 1, 2, (4, 5, 3), True, None, 4e-7),           # more ghost planes than the reach
    (configs.stencil27_op, 'zeros', 'backward', 1, 1, (3, 4, 3), True, None, 1e-14),
    (configs.heat3d_op, 'zeros', 'forward', 1, 1, (20, 17, 19), True, None, 4e-7),      # several z-chunks: rotated order
    (configs.heat3d_op, 'zeros', 'forward', 2, 2, (5, 4, 6), True, MarchTuning(exchange=True, ry=2, ty=6), 5e-7),
    (configs.heat3d_op, None, 'forward', 2, 2, (4, 5, 4), True, MarchTuning(exchange=False, ry=2, ty=4, sx=4), 5e-7),
])
def test_peer_halo_kernels_replay(make, bh, which, steps, g, sizes, masked, tuning, tol):
    shape = (sum(sizes), 14, 68)
    op = make(shape=shape, boundary_handling=bh)
    ir = op.forward_ast_gpu if which == 'forward' else op.backward_ast_gpu
    assigns = op.forward_assignments if which == 'forward' else op.backward_assignments
    fin, fout = ir.input_fields[0].name, ir.output_fields[0].name
    dt = ir.input_fields[0].dtype.numpy_dtype
    u = emu.aligned_empty(shape, dt)
    u[...] = np.random.default_rng(5).standard_normal(shape)
    ref = u.copy()
    for _ in range(steps):
        ref = evaluate(assigns, {fin: ref}, boundary_handling=bh)[fout].astype(dt)
    ins = _slabs(u, sizes, g)
    halo = max(ir.halo(fin)[0])
    world = len(sizes)
    for rank, n in enumerate(sizes):
        start, local_u = ins[rank]
        local_ir_op = make(shape=local_u.shape, boundary_handling=bh)
        local_ir = local_ir_op.forward_ast_gpu if which == 'forward' else local_ir_op.backward_ast_gpu
        ek = emit_march_chain(local_ir, tuning, peer=True) if steps == 2 else emit_march(local_ir, tuning, masked=masked, peer=True)
        assert ek.plan['peer'] == 1 and ek.name.endswith('_peer') and '#define PSAD_PEER 1' in ek.source
        local_out = emu.aligned_empty(local_u.shape, dt, np.nan)
        # what _PeerHalo.run launches: ONE range over all owned planes
        whole = slab_ranges(shape, start, n, g, False, False, 'zeros' if bh == 'zeros' else 'none', ir.ghost_layers, 3,
                            steps, halo if steps > 1 else None)[0]
        flags = np.array([7, 7, 0], dtype=np.uint32)
        lo = ins[rank - 1][1] if rank > 0 else None
        hi = ins[rank + 1][1] if rank < world - 1 else None
        arrays = [local_out if f.name == fout else local_u for f in ek.fields]
        peers = lambda nb: None if nb is None else [nb for _ in ek.fields]     # noqa: E731 (outputs are never staged)
        emu.run(ek, arrays, launch_range=whole, full=True, sm_count=2,
                peer=dict(lo=peers(lo), hi=peers(hi), ghost_planes=g, flags=flags, expect=7))
        assert flags[2] == 0
        got = local_out[g:g + n]
        assert not np.isnan(got).any(), 'rank %d read its own (stale) ghost planes' % rank
        np.testing.assert_allclose(got, ref[start:start + n], rtol=0, atol=tol)
        assert np.isnan(local_out[:g]).all() and np.isnan(local_out[g + n:]).all()      # ghost planes are not written
        # a neighbour that has not finished its previous launch: the wait is reached (recorded here, spun on by the device)
        if lo is not None or hi is not None:
            flags[:] = (6, 6, 0)
            emu.run(ek, arrays, launch_range=whole, full=True, sm_count=2,
                    peer=dict(lo=peers(lo), hi=peers(hi), ghost_planes=g, flags=flags, expect=7))
            assert flags[2] == 1


def test_peer_plan_is_validated():
    from pystencils_autodiff_b200 import runtime
    shape = (8, 14, 68)
    op = configs.heat3d_op(shape=shape, boundary_handling='zeros')
    u = emu.aligned_empty(shape, np.float32, 0.0)
    out = emu.aligned_empty(shape, np.float32, 0.0)
    flags = np.zeros(3, dtype=np.uint32)
    ek = emit_march(op.forward_ast_gpu, None, masked=True, peer=True)
    arrays = [out if f.name == 'out' else u for f in ek.fields]
    with pytest.raises(RuntimeError, match='at least one ghost plane'):
        emu.run(ek, arrays, full=True, peer=dict(lo=arrays, hi=None, ghost_planes=0, flags=flags, expect=0))
    with pytest.raises(RuntimeError, match='too few planes'):
        thin = [emu.aligned_empty((2,) + shape[1:], np.float32, 0.0)] * 2
        emu.run(ek, arrays, full=True, peer=dict(lo=thin, hi=None, ghost_planes=1, flags=flags, expect=0))
    # a kernel built without peer halos refuses a peer description and vice versa
    plain = emit_march(op.forward_ast_gpu, None, masked=True)
    with pytest.raises(RuntimeError, match='not a peer-halo kernel'):
        emu.run(plain, arrays, full=True, peer=dict(lo=arrays, hi=None, ghost_planes=1, flags=flags, expect=0))
    plan = runtime.make_plan(ek.plan)
    assert plan.reserved[2] == 1


def test_peer_launches_start_in_the_middle_of_the_slab():
    """psad_plan_launch_peer rotates the z-chunk order by half: the chunks that touch ghost planes (a wait for the
    neighbour, loads over NVLink) are not the first work of every CTA."""
    import ctypes
    import struct
    from pystencils_autodiff_b200 import runtime
    shape = (66, 64, 128)
    op = configs.heat3d_op(shape=shape, boundary_handling='zeros')
    ek = emit_march(op.forward_ast_gpu, None, masked=False, peer=True)
    L = runtime.lib()
    L.psad_args_size.restype = ctypes.c_size_t
    nbytes = int(L.psad_args_size())
    plan = runtime.make_plan(ek.plan)
    fa = (runtime.FieldArg * 2)()
    for i in range(2):
        fa[i].ptr = 4096
        fa[i].shape[:] = shape
        fa[i].stride[:] = [shape[1] * shape[2], shape[2], 1, 0]
    rng = runtime.Range()
    for d in range(3):
        rng.iter_lo[d] = rng.write_lo[d] = 0
        rng.iter_hi[d] = rng.write_hi[d] = shape[d]
    rng.write_lo[0], rng.write_hi[0] = 1, shape[0] - 1
    P = runtime.Peer()
    flags = np.zeros(4, dtype=np.uint32)
    for i in range(2):
        P.lo_ptr[i] = P.hi_ptr[i] = 8192
    P.lo_planes = P.hi_planes = shape[0]
    P.flag_lo = P.flag_hi = flags.ctypes.data
    P.error_flag = flags.ctypes.data + 4
    P.ghost_planes = 1
    args = ctypes.create_string_buffer(nbytes)
    grid = (ctypes.c_uint * 3)()
    L.psad_plan_launch_peer.argtypes = [ctypes.POINTER(runtime.Plan), ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(runtime.FieldArg), ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                        ctypes.c_int, ctypes.POINTER(runtime.Range), ctypes.POINTER(runtime.Peer),
                                        ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint)]
    rc = L.psad_plan_launch_peer(ctypes.byref(plan), 4, 1, fa, 2, None, 0, ctypes.byref(rng), ctypes.byref(P), args, nbytes, grid)
    assert rc == 0, L.psad_last_error()
    tiles_x, tiles_y, n_chunks, chunk = struct.unpack_from('4i', args.raw, 736)
    lo_end, hi_begin, lo_shift, hi_shift = struct.unpack_from('4i', args.raw, 780)
    rot, = struct.unpack_from('i', args.raw, 796)
    assert n_chunks >= 2 and rot == n_chunks // 2
    assert (lo_end, hi_begin, lo_shift, hi_shift) == (1, shape[0] - 1, shape[0] - 2, shape[0] - 2)
