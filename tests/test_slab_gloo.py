"""Slab decomposition + ghost-plane exchange on CPU: world_size 2 (and 3), gloo backend.

The kernels themselves need a GPU; here the *host logic* (decomposition, ranges, exchange, global-boundary
handling) is driven with the oracle as the per-slab evaluator: exchange ghosts, evaluate the local padded array,
keep the owned planes, compare with the oracle on the global array."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _collect(procs, q, world, timeout=600):
    """Results of all workers; fails as soon as one of them has died instead of waiting for the queue to time out."""
    import queue as _queue
    import time
    results, t0 = [], time.time()
    while len(results) < world:
        try:
            results.append(q.get(timeout=2))
        except _queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() - t0 > timeout:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                raise AssertionError('worker exit codes %s after %.0f s' % ([p.exitcode for p in procs], time.time() - t0))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return results


def _worker(rank, world, port, name, bh, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import evaluate
        from pystencils_autodiff_b200.configs import make_config
        from pystencils_autodiff_b200.datahandling import SlabDataHandling
        gshape = {'c2': (12, 9), 'c3': (12, 5, 6), 'c4': (9, 5, 6)}[name]
        op_g = make_config(name, shape=gshape, dtype='float64', boundary_handling=bh)
        g = max(op_g.forward_ast_gpu.max_halo[0])
        dh = SlabDataHandling(gshape, rank, world, g, device='cpu', backend='torch')
        rng = np.random.default_rng(5)
        glob_u = rng.normal(size=gshape)
        glob_go = rng.normal(size=gshape)
        dh.add_array('u', dtype=np.float64)
        dh.add_array('diffout', dtype=np.float64)
        sl = slice(dh.dec.start, dh.dec.start + dh.dec.n_local)
        dh.owned('u').copy_(torch.from_numpy(glob_u[sl]))
        dh.owned('diffout').copy_(torch.from_numpy(glob_go[sl]))
        dh.synchronization_function(['u', 'diffout'])()
        assert [c[0] for c in dh.call_queue] == ['Communication', 'Communication']
        # local evaluation on the padded slab (array ends act as zero boundary = never-received ghost planes)
        op_l = make_config(name, shape=dh.dec.local_shape, dtype='float64', boundary_handling=bh)
        res = {}
        for key, asg, src in (('out', op_l.forward_assignments, 'u'), ('diffu', op_l.backward_assignments, 'diffout')):
            local = evaluate(asg, {src: dh.gpu_arrays[src].numpy()}, 'zeros')[key][dh.dec.owned]
            ref = evaluate(getattr(op_g, 'forward_assignments' if key == 'out' else 'backward_assignments'),
                           {src: glob_u if src == 'u' else glob_go}, bh)[key][sl]
            if bh is None:
                # 'none': the kernel launch ranges clip the iteration space to the global interior
                ir = op_l.forward_ast_gpu if key == 'out' else op_l.backward_ast_gpu
                interior, lo, hi = dh.dec.ranges('none', ir.ghost_layers, ir.ndim)
                mask = np.zeros(dh.dec.local_shape, dtype=bool)
                for r in (interior, lo, hi):
                    if r is not None:
                        mask[tuple(slice(a, b) for a, b in zip(r['iter_lo'], r['iter_hi']))] = True
                local = np.where(mask[dh.dec.owned], local, 0.0)
            res[key] = float(np.abs(local - ref).max())
        # the write ranges of the three launches tile the owned planes exactly once
        ir = op_l.forward_ast_gpu
        cover = np.zeros(dh.dec.local_shape[0], dtype=int)
        for r in dh.dec.ranges(ir.boundary, ir.ghost_layers, ir.ndim):
            if r is not None:
                cover[r['write_lo'][0]:r['write_hi'][0]] += 1
        assert list(cover[dh.dec.owned]) == [1] * dh.dec.n_local and cover.sum() == dh.dec.n_local
        gathered = dh.gather_array('u')
        res['gather'] = float(np.abs(gathered - glob_u).max())
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name,bh,world', [('c3', 'zeros', 2), ('c3', None, 2), ('c2', 'zeros', 2), ('c4', 'zeros', 3),
                                           ('c4', None, 3)])
def test_sharded_equals_global(name, bh, world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + (hash((name, bh, world)) % 300)
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, bh, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(world))
    for r in range(world):
        for k, v in results[r].items():
            assert v < 1e-14, (r, k, v)


def test_decomposition_ranges():
    from pystencils_autodiff_b200.datahandling import SlabDecomposition
    d = SlabDecomposition((10, 8, 8), rank=1, world_size=3, ghost_layers=1)
    assert d.counts == [4, 3, 3] and d.start == 4 and d.local_shape == (5, 8, 8)
    assert (d.lo_rank, d.hi_rank) == (0, 2)
    interior, lo, hi = d.ranges('zeros', 0, 3)
    assert interior['write_lo'][0] == 2 and interior['write_hi'][0] == 3
    assert (lo['write_lo'][0], lo['write_hi'][0]) == (1, 2) and (hi['write_lo'][0], hi['write_hi'][0]) == (3, 4)
    first = SlabDecomposition((10, 8, 8), rank=0, world_size=3, ghost_layers=1)
    interior, lo, hi = first.ranges('none', 1, 3)
    assert lo is None and hi is not None
    assert interior['iter_lo'] == [2, 1, 1] and interior['write_lo'] == [1, 0, 0]   # global plane 0 is border: zero
    with pytest.raises(ValueError):
        SlabDecomposition((4, 8, 8), rank=0, world_size=4, ghost_layers=1)


def test_data_handling_registry_and_call_queue_single_rank():
    """SlabDataHandling's registry vocabulary on one rank (CPU tensors): arrays, fill, swap, host mirrors, queue."""
    from pystencils_autodiff_b200.datahandling import SlabDataHandling
    dh = SlabDataHandling((6, 8), 0, 1, 1, device='cpu', backend='torch')
    u, v = dh.add_arrays('u, v', dtype=np.float64)
    assert u.name == 'u' and dh.fields['v'].shape == (8, 8) and dh.gpu_arrays['u'].shape == (8, 8)   # 6 + 2 ghost rows
    w = dh.add_array_like('w', 'u')
    assert w.dtype.numpy_dtype == np.float64
    dh.fill('u', 3.0)
    assert float(dh.owned('u').sum()) == 3.0 * 6 * 8 and float(dh.gpu_arrays['u'].sum()) == 3.0 * 6 * 8
    dh.swap('u', 'v')
    assert float(dh.owned('v').sum()) == 3.0 * 48 and float(dh.owned('u').sum()) == 0
    host = dh.to_cpu('v')
    host += 1
    dh.to_gpu('v')
    assert float(dh.owned('v')[0, 0]) == 4.0
    dh.synchronization_function(['u'])()          # single rank: records the marker, moves nothing
    assert [c[0] for c in dh.call_queue] == ['Fill', 'Swap', 'DataTransfer', 'DataTransfer', 'Communication']
    assert np.array_equal(dh.gather_array('v'), dh.owned('v').numpy())
    with pytest.raises(ValueError):
        dh.add_array('u')


def _replay_kernel_class():
    """The shared ``ReplayKernel`` (tests/replay_kernels.py: the launch is a CPU replay of the emitted kernel) and the
    emulator module."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import march_emulator as emu
    from replay_kernels import ReplayKernel
    return ReplayKernel, emu


def _steps_worker(rank, world, port, bh, steps, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from pystencils_autodiff_b200 import configs
        from pystencils_autodiff_b200.datahandling import SlabDataHandling
        ReplayKernel, emu = _replay_kernel_class()
        gshape = (6 * world, 20, 132)
        rng = np.random.default_rng(3)
        glob_u = rng.standard_normal(gshape).astype(np.float32)
        res = {}
        for fuse in (False, True):
            dh = SlabDataHandling(gshape, rank, world, 2, device='cpu', backend='torch')   # 2 ghost planes = 2 x halo
            dh.add_arrays('u, out', dtype=np.float32)
            op_l = configs.heat3d_op(shape=dh.dec.local_shape, boundary_handling=bh)
            kernel = ReplayKernel(op_l.forward_ast_gpu)
            sl = slice(dh.dec.start, dh.dec.start + dh.dec.n_local)
            dh.owned('u').copy_(torch.from_numpy(glob_u[sl]))
            del ReplayKernel.launches[:]
            dh.run_steps(kernel, steps, fuse=fuse)
            res[fuse] = dh.gather_array('u')
            kinds = [c[0] for c in dh.call_queue]
            n_launch = (steps // 2 + steps % 2) if fuse else steps
            assert kinds.count('Communication') == n_launch and kinds.count('Swap') == n_launch
            parts = 1 + int(rank > 0) + int(rank < world - 1)     # interior + the planes next to each neighbour
            assert sum('x2' in n for n in ReplayKernel.launches) == parts * (steps // 2 if fuse else 0)
            assert len(ReplayKernel.launches) == parts * n_launch
            # the same loop through the reference's TimeLoop API (add_call + swap): fused pairs by default on a data
            # handling that stores 2 x halo ghost planes, single steps with fuse_steps=False — bit-identical to run_steps
            dh.owned('u').copy_(torch.from_numpy(glob_u[sl]))
            del dh.call_queue[:]
            tl = dh.create_timeloop(fuse_steps=None if fuse else False)
            tl.add_call(kernel, {'halo_fields': ['u']})
            tl.swap('u', 'out')
            tl.run(steps)
            assert tl.fused_last_run == fuse and [c[0] for c in dh.call_queue] == ['TimeloopRun']
            assert np.array_equal(dh.gather_array('u'), res[fuse])
        q.put((rank, res[False], res[True]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('bh,world,steps', [('zeros', 2, 5), (None, 3, 4)])
def test_run_steps_fused_on_slabs(bh, world, steps):
    """``SlabDataHandling.run_steps``: single-step launches with one-plane exchanges and fused pairs with one two-plane
    exchange per pair both reproduce the unsharded time loop (oracle); the fused slab launches equal the unsharded
    fused launches bit for bit.  The kernels are the emitted march / fused-step kernels, replayed on the CPU."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29950 + world
    procs = [ctx.Process(target=_steps_worker, args=(r, world, port, bh, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = _collect(procs, q, world)
    sys.path.insert(0, ROOT)
    from oracle import evaluate
    from pystencils_autodiff_b200 import configs
    from pystencils_autodiff_b200.emit_chain import emit_march_chain
    from pystencils_autodiff_b200.emit import emit_march
    _, emu = _replay_kernel_class()
    gshape = (6 * world, 20, 132)
    glob_u = np.random.default_rng(3).standard_normal(gshape).astype(np.float32)
    op = configs.heat3d_op(shape=gshape, boundary_handling=bh)
    ref = glob_u.astype(np.float64)
    for _ in range(steps):
        ref = evaluate(op.forward_assignments, {'u': ref}, boundary_handling=bh)['out']
    # unsharded launches of the same kernels: pairs, then the odd step
    x2 = emit_march_chain(op.forward_ast_gpu)
    x1 = emit_march(op.forward_ast_gpu, None, masked=bh != 'zeros')
    a = emu.aligned_empty(gshape, np.float32)
    a[...] = glob_u
    b = emu.aligned_empty(gshape, np.float32, 0.0)
    for n in [2] * (steps // 2) + [1] * (steps % 2):
        emu.run(x2 if n == 2 else x1, [b, a])
        a, b = b, a
    for rank, single, fused in results:
        assert np.abs(single - ref).max() < 5e-6, rank
        assert np.abs(fused - ref).max() < 5e-6, rank
        assert np.array_equal(fused, a), rank


def test_data_handling_queue_vocabulary(tmp_path):
    """require_autograd, extract_tensor, save_fields and merge_swaps_with_kernel_calls (graph_datahandling.py:329-355,
    framework_integration/datahandling.py:190-200) on CPU tensors."""
    from pystencils_autodiff_b200.datahandling import SlabDataHandling
    dh = SlabDataHandling((6, 8), 0, 1, 1, device='cpu', backend='torch')
    dh.add_arrays('u, v', dtype=np.float64)
    dh.fill('u', 2.0)
    dh.require_autograd(True, 'u', 'nonexistent')
    assert dh.gpu_arrays['u'].requires_grad and not dh.gpu_arrays['v'].requires_grad
    dh.require_autograd(False, 'u')
    assert dh.extract_tensor('u').shape == (6, 8) and dh.extract_tensor(dh.fields['u'], with_ghost_layers=True).shape == (8, 8)
    dh.save_fields(['u', 'v'], str(tmp_path / 'snap'))
    saved = np.load(str(tmp_path / 'snap') + '.rank0.npz')
    assert saved['u'].shape == (6, 8) and float(saved['u'].sum()) == 96.0
    assert 'FieldOutput' in str(dh) and "('Fill', 'u')" in str(dh)
    queue = [('KernelCall', 'k1'), ('Swap', 'u', 'v'), ('Communication', 'u', None, True), ('Swap', 'u', 'v'),
             ('KernelCall', 'k2', 2), ('Swap', 'a', 'b'), ('Swap', 'u', 'v'), ('Fill', 'u')]
    merged = dh.merge_swaps_with_kernel_calls(queue)
    assert merged == [('KernelCall+Swap', 'k1', (('u', 'v'),)), ('Communication', 'u', None, True), ('Swap', 'u', 'v'),
                      ('KernelCall+Swap', 'k2', 2, (('a', 'b'), ('u', 'v'))), ('Fill', 'u')]
    assert dh.merge_swaps_with_kernel_calls() is dh.call_queue          # the recorded queue, merged in place


def test_arrays_with_their_own_shape_are_replicated():
    """MultiShapeDatahandling vocabulary (framework_integration/datahandling.py:54-132): ``spatial_shape`` / the
    description syntax give arrays of their own size — whole on every rank, no ghost planes, run unsharded."""
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    from pystencils_autodiff_b200.datahandling import SlabDataHandling
    dh = SlabDataHandling((6, 8), 0, 1, 1, device='cpu', backend='torch')
    u, = dh.add_arrays('u', dtype=np.float64)
    x, y = dh.add_arrays('x, y(2): float32[20,30]')
    w = dh.add_array('w', spatial_shape=(5, 4))
    assert dh.gpu_arrays['u'].shape == (8, 8) and dh.gpu_arrays['x'].shape == (20, 30)
    assert dh.gpu_arrays['y'].shape == (20, 30, 2) and y.index_dimensions == 1 and dh.gpu_arrays['x'].dtype == torch.float32
    assert w.spatial_shape == (5, 4) and dh.owned('w').shape == (5, 4) and dh.owned('u').shape == (6, 8)
    dh.fill('x', 1.5)
    assert float(dh.gpu_arrays['x'].sum()) == 1.5 * 600 and dh.gather_array('x').shape == (20, 30)
    assert dh.extract_tensor('w').shape == (5, 4)
    dh.synchronization_function(['x'])()           # never exchanged
    # a kernel over replicated arrays is called on the whole arrays, without a launch range; mixing is an error
    import pystencils_autodiff_b200 as ps
    a, b = ps.fields('x, z: float32[20,30]')
    dh.add_array('z', spatial_shape=(20, 30))
    seen = {}

    class Probe(CompiledKernel):
        def __call__(self, *, _range=None, _variant=None, _stream=None, **kw):
            seen.update(range=_range, names=sorted(kw))

    op = ps.AutoDiffOp([ps.Assignment(b.center, 2 * a[0, 1])], op_name='rep', boundary_handling='zeros')
    dh.run_kernel(Probe(op.forward_ast_gpu))
    assert seen == dict(range=None, names=['x', 'z'])
    c, d = ps.fields('u, z: float64[8,8]')
    mixed = ps.AutoDiffOp([ps.Assignment(d.center, c[0, 0])], op_name='mixed')
    with pytest.raises(ValueError, match='mixes'):
        dh.run_kernel(Probe(mixed.forward_ast_gpu))


def test_run_kernel_accepts_the_op_like_the_reference_test():
    """tests/test_datahandling.py:17-35: ``dh.run_kernel(op, a=3)`` with the torch op itself; here the op's forward kernel
    runs on the registered arrays (CPU replay of the emitted generic kernel as the launch) and writes ``z`` in place.
    Plain callables get every registered array by name (single rank only)."""
    import sympy
    import pystencils_autodiff_b200 as ps
    from oracle.evaluate import evaluate
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    from pystencils_autodiff_b200.datahandling import SlabDataHandling
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import march_emulator as emu

    class GenericReplay(CompiledKernel):
        def __call__(self, *, _range=None, _variant=None, _stream=None, **kw):
            ek = self._emitted['generic']
            emu.run_generic(ek, [kw[f.name].numpy() for f in ek.fields], [float(kw[s_]) for s_ in self.scalars],
                            launch_range=_range)

    dh = SlabDataHandling((20, 30), 0, 1, 0, device='cpu', backend='torch')
    for n in 'xyz':
        dh.add_array(n)
    a = sympy.Symbol('a')
    z, y, x = ps.fields('z, y, x: float32[20,30]')
    op = ps.AutoDiffOp(ps.AssignmentCollection({z[0, 0]: x[0, 0] * sympy.log(a * x[0, 0] * y[0, 0])}), op_name='dhop')
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    fn.forward_kernel = GenericReplay(op.forward_ast_gpu)
    rng = np.random.default_rng(0)
    X, Y = rng.uniform(0.5, 1.5, (20, 30)).astype(np.float32), rng.uniform(0.5, 1.5, (20, 30)).astype(np.float32)
    dh.owned('x').copy_(torch.from_numpy(X))
    dh.owned('y').copy_(torch.from_numpy(Y))
    dh.run_kernel(fn, a=3)
    ref = evaluate(op.forward_assignments, {'x': X.astype(np.float64), 'y': Y.astype(np.float64)}, None, scalars={'a': 3.0})['z']
    np.testing.assert_allclose(dh.owned('z').numpy(), ref, rtol=0, atol=2e-6)
    seen = {}
    dh.run_kernel(lambda **kw: seen.update(kw), b=1)
    assert sorted(seen) == ['b', 'x', 'y', 'z'] and seen['x'].shape == (20, 30)
    with pytest.raises(TypeError):
        dh.run_kernel(42)


def test_reference_constructor_aliases():
    """``GraphDataHandling`` / ``PyTorchDataHandling`` with the reference's constructor arguments (single process)."""
    from pystencils_autodiff_b200.datahandling import GraphDataHandling, PyTorchDataHandling
    dh = PyTorchDataHandling((20, 30), device='cpu')
    assert dh.dec.world_size == 1 and dh.dec.g == 0 and dh.add_array('x').spatial_shape == (20, 30)
    dh = GraphDataHandling((10, 15), default_ghost_layers=1, periodicity=False, default_target='gpu', device='cpu')
    assert dh.dec.local_shape == (12, 15) and dh.create_timeloop(use_cuda_graph=False).time_steps_run == 0
    with pytest.raises(NotImplementedError):
        GraphDataHandling((10, 15), periodicity=True, device='cpu')
    with pytest.raises(NotImplementedError):
        GraphDataHandling((10, 15), default_target='cpu', device='cpu')


def _autograd_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import sympy as sp
        import pystencils_autodiff_b200 as ps
        from oracle import evaluate, forward_backward
        from pystencils_autodiff_b200 import configs
        from pystencils_autodiff_b200.datahandling import SlabDataHandling, create_slab_autograd_function
        from replay_kernels import ReplayKernel
        res = {}
        # ---- A: two chained steps of the 7-point stencil (fp32, march kernels), loss = sum(out2 * r) ----------------------
        gshape = (6 * world + 1, 10, 132)                    # uneven slabs: rank 0 owns one plane more
        rng = np.random.default_rng(8)
        U = rng.standard_normal(gshape).astype(np.float32)
        R = rng.standard_normal(gshape).astype(np.float32)
        dh = SlabDataHandling(gshape, rank, world, 1, device='cpu', backend='torch')
        n = dh.dec.n_local
        sl = slice(dh.dec.start, dh.dec.start + n)
        op_l = configs.heat3d_op(shape=(n,) + gshape[1:], boundary_handling='zeros')
        Step = create_slab_autograd_function(op_l, dh, kernel_class=ReplayKernel)
        u = torch.from_numpy(U[sl].copy()).requires_grad_(True)
        del ReplayKernel.launches[:]
        (o1,) = Step.apply(u)
        (o2,) = Step.apply(o1)
        assert o1._base is not None and o1._base.shape[0] == n + 2          # outputs are views of padded buffers
        (o2 * torch.from_numpy(R[sl])).sum().backward()
        parts = 1 + int(rank > 0) + int(rank < world - 1)          # interior + the planes next to each neighbour
        assert all('march' in name for name in ReplayKernel.launches) and len(ReplayKernel.launches) == 4 * parts
        op_g = configs.heat3d_op(shape=gshape, boundary_handling='zeros')
        r1 = evaluate(op_g.forward_assignments, {'u': U.astype(np.float64)}, 'zeros')['out']
        r2 = evaluate(op_g.forward_assignments, {'u': r1}, 'zeros')['out']
        d2 = evaluate(op_g.backward_assignments, {'diffout': R.astype(np.float64)}, 'zeros')['diffu']
        d1 = evaluate(op_g.backward_assignments, {'diffout': d2}, 'zeros')['diffu']
        res['chain_out'] = float(np.abs(o2.detach().numpy() - r2[sl]).max())
        res['chain_grad'] = float(np.abs(u.grad.numpy() - d1[sl]).max())
        kinds = [c[0] for c in dh.call_queue]
        assert kinds.count('Communication') == 4 and kinds.count('KernelCall') == 4
        # ---- A': five unrolled steps as ONE Function (single launches, then fused pairs with two-plane exchanges) ---------
        from pystencils_autodiff_b200.datahandling import create_slab_unrolled_function
        dh5 = SlabDataHandling(gshape, rank, world, 2, device='cpu', backend='torch')
        ref, dref = U.astype(np.float64), R.astype(np.float64)
        for _ in range(5):
            ref = evaluate(op_g.forward_assignments, {'u': ref}, 'zeros')['out']
            dref = evaluate(op_g.backward_assignments, {'diffout': dref}, 'zeros')['diffu']
        for fuse in (False, True):
            Five = create_slab_unrolled_function(op_l, dh5, 5, fuse=fuse, kernel_class=ReplayKernel)
            assert Five.launches == ([2, 2, 1] if fuse else [1] * 5)
            u5 = torch.from_numpy(U[sl].copy()).requires_grad_(True)
            n_comm = sum(1 for c in dh5.call_queue if c[0] == 'Communication')
            (o5,) = Five.apply(u5)
            (o5 * torch.from_numpy(R[sl])).sum().backward()
            assert sum(1 for c in dh5.call_queue if c[0] == 'Communication') - n_comm == 2 * len(Five.launches)
            assert torch.equal(u5.detach(), torch.from_numpy(U[sl]))                 # the input is never written
            res['x5_out_%s' % fuse] = float(np.abs(o5.detach().numpy() - ref[sl]).max())
            res['x5_grad_%s' % fuse] = float(np.abs(u5.grad.numpy() - dref[sl]).max())
        with pytest.raises(ValueError, match='ghost layers'):
            create_slab_unrolled_function(op_l, dh, 4, fuse=True, kernel_class=ReplayKernel)     # dh stores one ghost plane
        # ---- B: two inputs, non-linear, offsets along dim 0 on both, one constant field (fp64, generic kernels) ---------
        gshape = (5 * world + 1, 6, 7)
        dh2 = SlabDataHandling(gshape, rank, world, 1, device='cpu', backend='torch')
        n = dh2.dec.n_local
        sl = slice(dh2.dec.start, dh2.dec.start + n)

        def make(shape, consts=True):
            a, b, c, out = ps.fields('a, b, c, out: float64[%d,%d,%d]' % shape)
            asg = ps.AssignmentCollection({out.center: a[1, 0, 0] * b[0, 0, 0] + sp.sin(a[0, -1, 0]) + 0.3 * b[-1, 0, 1] * c[0, 0, 0]})
            return ps.AutoDiffOp(asg, op_name='nl', boundary_handling='zeros', constant_fields=[c] if consts else [])
        op_l, op_g = make((n,) + gshape[1:]), make(gshape)
        ins = {k: rng.uniform(0.5, 1.5, gshape) for k in 'abc'}
        G = rng.standard_normal(gshape)
        F = create_slab_autograd_function(op_l, dh2, kernel_class=ReplayKernel)
        tens = [torch.from_numpy(ins[f.name][sl].copy()).requires_grad_(f.name != 'c') for f in op_l.forward_input_fields]
        (out,) = F.apply(*tens)
        out.backward(torch.from_numpy(G[sl].copy()))
        ref_out, ref_grads = forward_backward(op_g, ins, {'out': G})
        res['nl_out'] = float(np.abs(out.detach().numpy() - ref_out['out'][sl]).max())
        for f, t in zip(op_l.forward_input_fields, tens):
            if f.name == 'c':
                assert t.grad is None
            else:
                res['nl_d' + f.name] = float(np.abs(t.grad.numpy() - ref_grads['diff' + f.name][sl]).max())
        with pytest.raises(ValueError, match='owned planes'):
            F.apply(*[torch.zeros((n + 1,) + gshape[1:], dtype=torch.float64)] * 3)
        # ---- C: scalar input, vector output (index dimension): the curl-like case of tests/test_tfmad.py:341-401, 2-D ------
        gshape = (4 * world + 1, 10)
        dh3 = SlabDataHandling(gshape, rank, world, 1, device='cpu', backend='torch')
        n = dh3.dec.n_local
        sl = slice(dh3.dec.start, dh3.dec.start + n)

        def curl(shape):
            v = ps.Field.create_fixed_size('v', shape, index_dimensions=0, dtype=np.float64)
            c = ps.Field.create_fixed_size('c', shape + (2,), index_dimensions=1, dtype=np.float64)
            disc = ps.fd.Discretization2ndOrder(dx=1)
            return ps.AutoDiffOp(ps.AssignmentCollection([ps.Assignment(c.center(0), disc(ps.fd.Diff(v, 0))),
                                                          ps.Assignment(c.center(1), disc(ps.fd.Diff(v, 1)))], []),
                                 op_name='curl', boundary_handling='zeros')
        op_l, op_g = curl((n, gshape[1])), curl(gshape)
        V, GC = rng.standard_normal(gshape), rng.standard_normal(gshape + (2,))
        C = create_slab_autograd_function(op_l, dh3, kernel_class=ReplayKernel)
        v = torch.from_numpy(V[sl].copy()).requires_grad_(True)
        (c,) = C.apply(v)
        assert tuple(c.shape) == (n, gshape[1], 2)
        c.backward(torch.from_numpy(GC[sl].copy()))
        ref_out, ref_grads = forward_backward(op_g, {'v': V}, {'c': GC})
        res['curl_out'] = float(np.abs(c.detach().numpy() - ref_out['c'][sl]).max())
        res['curl_grad'] = float(np.abs(v.grad.numpy() - ref_grads['diffv'][sl]).max())
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_slab_autograd_function_two_ranks(world):
    """``create_slab_autograd_function`` (SURVEY.md §8e "Autograd"): forward = halo exchange + forward kernel on the owned
    planes, backward = the same exchange on the upstream gradient + the adjoint kernel; chained steps reuse the padded
    buffers.  Two gloo ranks, emitted kernels replayed on the CPU, compared with the oracle on the GLOBAL field."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_autograd_worker, args=(r, world, 29877 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(_collect(procs, q, world))
    for r in range(world):
        assert results[r]['chain_out'] < 2e-6 and results[r]['chain_grad'] < 2e-6, results[r]
        for k in ('nl_out', 'nl_da', 'nl_db', 'curl_out', 'curl_grad'):
            assert results[r][k] < 1e-12, (r, k, results[r])
        for k in ('x5_out_False', 'x5_grad_False', 'x5_out_True', 'x5_grad_True'):
            assert results[r][k] < 5e-6, (r, k, results[r])


def _e2e_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from fake_cuda import fake_cuda
        from oracle import evaluate
        from pystencils_autodiff_b200 import configs
        from pystencils_autodiff_b200.datahandling import SlabStencilOp
        from replay_kernels import ReplayKernel
        local = (6, 10, 132)
        gshape = (local[0] * world,) + local[1:]
        op = configs.heat3d_op(shape=local, boundary_handling='zeros')
        with fake_cuda():
            slab = SlabStencilOp(op, local, rank, world, device='cpu', backend='torch')
            slab.fwd, slab.bwd = ReplayKernel(op.forward_ast_gpu), ReplayKernel(op.backward_ast_gpu)
            g0 = torch.Generator().manual_seed(5 + rank)
            slab.randomize(g0)
            r = slab.end_to_end(2, dist.barrier)
            plane = int(np.prod(local[1:])) * 4
            cells = int(np.prod(local))
            # every input plane once, plus the g = 1 planes next to each side uploaded ahead for the neighbour exchange
            assert r['ms_per_step'] > 0 and r['h2d'] == 2 * 4 * cells + 2 * 2 * plane and r['d2h'] == 2 * 4 * cells
            assert r['matches_resident'] is True and r['copy_only_ms'] > 0
            assert 0 in r['checked_planes'] and local[0] - 1 in r['checked_planes']
            # the host arrays the streamed operator filled, gathered over the ranks
            got = {}
            for n in ('u', 'out', 'diffout', 'diffu'):
                parts = [torch.empty(local, dtype=torch.float32) for _ in range(world)]
                dist.all_gather(parts, slab._host[n].contiguous())
                got[n] = torch.cat(parts, 0).numpy()
        op_g = configs.heat3d_op(shape=gshape, boundary_handling='zeros')
        ref_out = evaluate(op_g.forward_assignments, {'u': got['u'].astype(np.float64)}, 'zeros')['out']
        ref_du = evaluate(op_g.backward_assignments, {'diffout': got['diffout'].astype(np.float64)}, 'zeros')['diffu']
        scale = max(np.abs(ref_out).max(), np.abs(ref_du).max(), 1e-300)
        assert np.abs(got['u']).max() > 0
        q.put((rank, float(np.abs(got['out'] - ref_out).max() / scale), float(np.abs(got['diffu'] - ref_du).max() / scale)))
    finally:
        dist.destroy_process_group()


def test_end_to_end_leg_with_host_buffers_two_ranks():
    """``SlabStencilOp.end_to_end`` at N > 1 (bench.py's ``e2e`` leg: every rank streams its slab from host memory in chunks,
    the planes next to a neighbouring rank exchanged between the "devices") executed on CPU tensors with stand-in streams
    and events (tests/fake_cuda.py) and replayed kernels, three gloo ranks (a middle rank has neighbours on both sides)."""
    world = 3
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_e2e_worker, args=(r, world, 29891, q)) for r in range(world)]
    for p in procs:
        p.start()
    for rank, e_out, e_du in _collect(procs, q, world):
        assert e_out < 1e-6 and e_du < 1e-6, (rank, e_out, e_du)


def test_slab_ranges_cover_the_owned_planes_exactly_once():
    """Pure index logic over many decompositions: the interior / lo / hi write ranges tile the owned planes exactly once
    (single steps and fused pairs), iteration ranges stay inside the local array, and — fused pairs — contain every
    written plane that belongs to the global iteration space."""
    import itertools
    from pystencils_autodiff_b200.datahandling import slab_ranges
    checked = 0
    for N, world, halo, steps, boundary in itertools.product([12, 17, 40], [1, 2, 3, 4], [1, 2], [1, 2], ['zeros', 'none']):
        g = steps * halo
        base, rem = divmod(N, world)
        counts = [base + (1 if r < rem else 0) for r in range(world)]
        if world > 1 and min(counts) < 2 * g:
            continue
        for rank in range(world):
            start, n = sum(counts[:rank]), counts[rank]
            parts = slab_ranges((N, 9, 11), start, n, g, rank > 0, rank < world - 1, boundary, halo, 3,
                                steps=steps, halo=halo if steps > 1 else None)
            cover = np.zeros(n + 2 * g, dtype=int)
            for p in parts:
                if p is None:
                    continue
                cover[p['write_lo'][0]:p['write_hi'][0]] += 1
                assert p['write_lo'][1:] == [0, 0] and p['write_hi'][1:] == [9, 11]
                assert 0 <= p['iter_lo'][0] <= p['iter_hi'][0] <= n + 2 * g
                if steps > 1:
                    dom_lo, dom_hi = (halo, N - halo) if boundary == 'none' else (0, N)
                    for z in range(p['write_lo'][0], p['write_hi'][0]):
                        inside = dom_lo <= z - g + start < dom_hi
                        assert (p['iter_lo'][0] <= z < p['iter_hi'][0]) == inside
            assert list(cover) == [0] * g + [1] * n + [0] * g, (N, world, halo, steps, boundary, rank)
            interior = parts[0]
            if interior is not None and steps > 1 and world > 1:
                # planes of the interior launch depend on no ghost plane
                lo_dep = interior['write_lo'][0] - steps * halo
                hi_dep = interior['write_hi'][0] + steps * halo
                assert lo_dep >= (g if rank > 0 else 0) and hi_dep <= n + g + (0 if rank < world - 1 else g)
            checked += 1
    assert checked > 100


def test_computation_graph_of_a_recorded_time_loop(tmp_path):
    """``ComputationGraph`` over the queue a ``SlabDataHandling`` records (tests/test_graph_datahandling.py:77-88 builds the
    reference's from ``call_queue`` and writes a dot file): array versions, read / write maps, levels of independent calls."""
    from pystencils_autodiff_b200 import configs
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    from pystencils_autodiff_b200.computationgraph import ComputationGraph
    from pystencils_autodiff_b200.datahandling import SlabDataHandling

    class Silent(CompiledKernel):
        def __call__(self, **kw):
            pass

    dh = SlabDataHandling((8, 10, 12), 0, 1, 1, device='cpu', backend='torch')
    dh.add_arrays('u, out, diffout, diffu')
    op = configs.heat3d_op(shape=dh.dec.local_shape)
    fwd, bwd = Silent(op.forward_ast_gpu), Silent(op.backward_ast_gpu)
    dh.run_kernel(fwd, halo_fields=['u'])
    dh.run_kernel(bwd, halo_fields=['diffout'])          # independent of the forward kernel
    dh.swap('u', 'out')
    dh.run_kernel(fwd, halo_fields=['u'])
    g = ComputationGraph(dh)
    assert dh.kernel_io['heat3d_forward_gpu'] == (['u'], ['out']) and dh.kernel_io['heat3d_backward_gpu'] == (['diffout'], ['diffu'])
    assert [n.kind for n in g.computation_nodes] == ['communication', 'kernel', 'communication', 'kernel', 'swap',
                                                     'communication', 'kernel']
    assert 'out #1' in g.writes and g.writes['out #1'].label == 'heat3d_forward_gpu'
    assert [n.label for n in g.reads['u #1']] == ['heat3d_forward_gpu', 'Swap u <-> out']
    levels = g.levels()
    assert {n.label for n in levels[0]} == {'ghost planes of u', 'ghost planes of diffout'}
    assert {n.label for n in levels[1]} == {'heat3d_forward_gpu', 'heat3d_backward_gpu'}      # may run concurrently
    assert [n.kind for n in levels[2]] == ['swap'] and [n.kind for n in levels[-1]] == ['kernel']
    dot = g.to_dot()
    assert dot.startswith('digraph') and '"u #0" -> ' in dot and 'heat3d_backward_gpu' in dot
    g.to_dot_file(str(tmp_path / 'graph.dot'), with_code=False)
    assert (tmp_path / 'graph.dot').read_text() == dot
    merged = ComputationGraph(dh.merge_swaps_with_kernel_calls(list(dh.call_queue)), dh.kernel_io)
    assert len(merged.computation_nodes) == len(g.computation_nodes)
    with pytest.raises(KeyError):
        ComputationGraph([('KernelCall', 'unknown_kernel')])


def _periodic_reference(assigns, key, src_name, glob, g):
    """Global field periodic along dim 0, 'zeros' along the other axes: evaluate on the field wrapped by ``g`` planes."""
    from oracle import evaluate
    wrapped = np.concatenate([glob[-g:], glob, glob[:g]], axis=0)
    return evaluate(assigns, {src_name: wrapped}, 'zeros')[key][g:-g]


def _periodic_check(dh, op_g, gshape, glob_u, glob_go, bh):
    from oracle import evaluate
    from pystencils_autodiff_b200.configs import make_config
    g = dh.dec.g
    sl = slice(dh.dec.start, dh.dec.start + dh.dec.n_local)
    dh.owned('u').copy_(torch.from_numpy(glob_u[sl]))
    dh.owned('diffout').copy_(torch.from_numpy(glob_go[sl]))
    dh.synchronization_function(['u', 'diffout'])()
    name = op_g.op_name
    op_l = {'heat3d': 'c3', 'stencil27': 'c4'}[name]
    op_l = make_config(op_l, shape=dh.dec.local_shape, dtype='float64', boundary_handling='zeros')
    res = {}
    for key, asg_l, asg_g, src, glob in (('out', op_l.forward_assignments, op_g.forward_assignments, 'u', glob_u),
                                         ('diffu', op_l.backward_assignments, op_g.backward_assignments, 'diffout', glob_go)):
        local = evaluate(asg_l, {src: dh.gpu_arrays[src].numpy()}, 'zeros')[key][dh.dec.owned]
        ref = _periodic_reference(asg_g, key, src, glob, g)[sl]
        res[key] = float(np.abs(local - ref).max())
    # launch ranges: every rank has neighbours on both sides (or is its own), nothing is clipped along dim 0
    ir = op_l.forward_ast_gpu
    cover = np.zeros(dh.dec.local_shape[0], dtype=int)
    for r in dh.dec.ranges('none' if bh is None else 'zeros', 1, ir.ndim):
        if r is not None:
            cover[r['write_lo'][0]:r['write_hi'][0]] += 1
            assert r['iter_lo'][0] >= g and r['iter_hi'][0] <= g + dh.dec.n_local
            assert r['iter_lo'][0] == r['write_lo'][0] and r['iter_hi'][0] == r['write_hi'][0]
    assert list(cover[dh.dec.owned]) == [1] * dh.dec.n_local
    return res


def _periodic_worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from pystencils_autodiff_b200.configs import make_config
        from pystencils_autodiff_b200.datahandling import SlabDataHandling
        gshape = (12, 5, 6)
        op_g = make_config(name, shape=(gshape[0] + 2,) + gshape[1:], dtype='float64', boundary_handling='zeros')
        dh = SlabDataHandling(gshape, rank, world, 1, device='cpu', backend='torch', periodic=True)
        assert dh.dec.lo_rank == (rank - 1) % world and dh.dec.hi_rank == (rank + 1) % world
        dh.add_array('u', dtype=np.float64)
        dh.add_array('diffout', dtype=np.float64)
        rng = np.random.default_rng(8)
        q.put((rank, _periodic_check(dh, op_g, gshape, rng.normal(size=gshape), rng.normal(size=gshape), 'zeros')))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name,world', [('c3', 2), ('c4', 3)])
def test_periodic_synchronisation_wraps_around(name, world):
    """graph_datahandling.py:305-316 -> pystencils' periodic ghost-layer copy: with ``periodic=True`` the first and last rank
    exchange planes (two ranks: both neighbours are the same peer, the receives are ordered accordingly)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_periodic_worker, args=(r, world, 29871 + world, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    for rank, res in _collect(procs, q, world):
        assert res['out'] < 1e-13 and res['diffu'] < 1e-13, (rank, res)


def test_periodic_single_rank_is_its_own_neighbour():
    from pystencils_autodiff_b200.configs import make_config
    from pystencils_autodiff_b200.datahandling import GraphDataHandling, SlabDataHandling
    gshape = (7, 5, 6)
    op_g = make_config('c3', shape=(gshape[0] + 2,) + gshape[1:], dtype='float64', boundary_handling='zeros')
    dh = SlabDataHandling(gshape, 0, 1, 1, device='cpu', backend='torch', periodic=True)
    dh.add_array('u', dtype=np.float64)
    dh.add_array('diffout', dtype=np.float64)
    rng = np.random.default_rng(2)
    res = _periodic_check(dh, op_g, gshape, rng.normal(size=gshape), rng.normal(size=gshape), 'zeros')
    assert res['out'] < 1e-13 and res['diffu'] < 1e-13
    # the reference's constructor: periodic along the decomposed axis is accepted, other axes are refused with a reason
    assert GraphDataHandling((8, 8), 1, periodicity=(True, False), device='cpu').dec.periodic
    with pytest.raises(NotImplementedError, match='dim 0 only'):
        GraphDataHandling((8, 8), 1, periodicity=True, device='cpu')


@pytest.mark.parametrize('name,shape,g,bh', [('c3', (10, 12, 136), 2, 'zeros'), ('c3', (9, 12, 132), 0, None),
                                             ('c3', (7, 8, 132), 1, 'zeros'), ('c4', (7, 8, 68), 2, 'zeros'),
                                             ('c2', (24, 136), 0, 'zeros')])
def test_timeloop_fuses_pairs_of_steps(name, shape, g, bh):
    """The reference's time-loop idiom — ``add_call(kernel)`` + ``swap(in, out)`` (graph_datahandling.py:152-197) — runs as
    fused pairs of steps (``out = S(S(u))``, one launch per two time steps) where ``run_steps`` would fuse: same result as
    ``run_steps`` bit for bit, the oracle chain within tolerance, ONE ``TimeloopRun`` record in the reference's vocabulary;
    ``fuse_steps=False`` keeps single steps.  Kernels replayed on the CPU from the emitted source."""
    from oracle import evaluate
    from pystencils_autodiff_b200.configs import make_config
    from pystencils_autodiff_b200.datahandling import SlabDataHandling
    from replay_kernels import ReplayKernel
    T = 5
    op = make_config(name, shape=shape, boundary_handling=bh)
    dt = op.forward_ast_gpu.input_fields[0].dtype.numpy_dtype        # c4: float64 (pairs are the default there too)
    U0 = np.random.default_rng(11).normal(size=shape).astype(dt)
    finals = {}
    for mode in (None, False, 'run_steps'):
        dh = SlabDataHandling(shape, 0, 1, g, device='cpu', backend='torch')
        dh.add_arrays('u, out', dtype=dt)
        kern = ReplayKernel(make_config(name, shape=dh.dec.local_shape, boundary_handling=bh).forward_ast_gpu)
        dh.owned('u').copy_(torch.from_numpy(U0))
        ReplayKernel.launches.clear()
        if mode == 'run_steps':
            dh.run_steps(kern, T)
        else:
            tl = dh.create_timeloop(fuse_steps=mode)
            tl.add_call(kern, {})
            tl.swap('u', 'out')
            assert (tl.fused_pair() is not None) == (mode is None)
            tl.run(T)
            assert tl.time_steps_run == T and tl.fused_last_run == (mode is None)
            kind, steps, recorded = dh.call_queue[-1]
            assert kind == 'TimeloopRun' and steps == T and [c[0] for c in recorded] == ['KernelCall', 'Swap']
            assert len(dh.call_queue) == 1
        fused_launches = [n for n in ReplayKernel.launches if n.endswith('_x2') or '_x2' in n]
        assert len(ReplayKernel.launches) == (T if mode is False else T // 2 + T % 2)
        assert len(fused_launches) == (0 if mode is False else T // 2)
        finals[mode] = dh.owned('u').clone()
    assert torch.equal(finals[None], finals['run_steps'])
    ref = U0.astype(np.float64)
    for _ in range(T):
        ref = evaluate(op.forward_assignments, dict(u=ref), bh)['out']
    for mode in (None, False):
        assert np.abs(finals[mode].numpy() - ref).max() <= (2e-6 if dt.itemsize == 4 else 1e-13) * np.abs(ref).max()


def test_timeloop_fused_pairs_are_refused_where_they_cannot_be_built():
    """``fuse_steps=True`` raises with the reason; the default silently keeps single steps."""
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    from pystencils_autodiff_b200.configs import make_config
    from pystencils_autodiff_b200.datahandling import SlabDataHandling
    dh = SlabDataHandling((4, 16, 132), 0, 1, 0, device='cpu', backend='torch')
    dh.add_arrays('u, f, g', dtype=np.float32)
    tv = CompiledKernel(make_config('c5', shape=(4, 16, 132)).forward_ast_gpu)       # two input fields
    for fuse, raises in ((None, False), (True, True)):
        tl = dh.create_timeloop(fuse_steps=fuse)
        tl.add_call(tv, {})
        tl.swap('u', 'g')
        if raises:
            with pytest.raises(ValueError, match='one-input'):
                tl.fused_pair()
        else:
            assert tl.fused_pair() is None
    heat = CompiledKernel(make_config('c3', shape=(4, 16, 132)).forward_ast_gpu)
    dh2 = SlabDataHandling((4, 16, 132), 0, 1, 0, device='cpu', backend='torch')
    dh2.add_arrays('u, out', dtype=np.float32)
    tl = dh2.create_timeloop(fuse_steps=True)
    tl.add_call(heat, {})
    tl.add_single_step_function(lambda: dh2.swap('u', 'out'))          # an opaque function: the loop cannot know it swaps
    with pytest.raises(ValueError, match='add_call'):
        tl.fused_pair()


def _periodic_steps_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import evaluate
        from pystencils_autodiff_b200.configs import make_config
        from pystencils_autodiff_b200.datahandling import SlabDataHandling
        ReplayKernel, _ = _replay_kernel_class()
        T = 5
        gshape = (6 * world, 6, 68)
        glob = np.random.default_rng(21).standard_normal(gshape).astype(np.float32)
        ref = glob.astype(np.float64)
        op_w = make_config('c3', shape=(gshape[0] + 2,) + gshape[1:], dtype='float64', boundary_handling='zeros')
        for _ in range(T):       # periodic along dim 0, 'zeros' along the other axes
            ref = evaluate(op_w.forward_assignments, {'u': np.concatenate([ref[-1:], ref, ref[:1]], 0)}, 'zeros')['out'][1:-1]
        res = {}
        for g, mode in ((1, 'steps'), (2, 'steps'), (2, 'loop')):
            dh = SlabDataHandling(gshape, rank, world, g, device='cpu', backend='torch', periodic=True)
            dh.add_arrays('u, out', dtype=np.float32)
            kernel = ReplayKernel(make_config('c3', shape=dh.dec.local_shape, boundary_handling='zeros').forward_ast_gpu)
            sl = slice(dh.dec.start, dh.dec.start + dh.dec.n_local)
            dh.owned('u').copy_(torch.from_numpy(glob[sl]))
            del ReplayKernel.launches[:]
            if mode == 'steps':
                dh.run_steps(kernel, T)              # default: pairs only where 2 x reach ghost planes are stored
                if g == 1:
                    with pytest.raises(ValueError, match='ghost planes'):
                        dh.run_kernel(kernel, halo_fields=['u'], fused_steps=2)
            else:
                tl = dh.create_timeloop()
                tl.add_call(kernel, {'halo_fields': ['u']})
                tl.swap('u', 'out')
                tl.run(T)
                assert tl.fused_last_run
            fused = sum('x2' in n for n in ReplayKernel.launches)
            assert (fused > 0) == (g == 2), (g, mode, ReplayKernel.launches)
            res['%d %s' % (g, mode)] = float(np.abs(dh.owned('u').numpy() - ref[sl]).max() / np.abs(ref).max())
        # autograd on the periodic slabs: two chained steps through the slab Function, loss = sum(out2 * r)
        from pystencils_autodiff_b200.datahandling import create_slab_autograd_function
        dh = SlabDataHandling(gshape, rank, world, 1, device='cpu', backend='torch', periodic=True)
        sl = slice(dh.dec.start, dh.dec.start + dh.dec.n_local)
        op_l = make_config('c3', shape=(dh.dec.n_local,) + gshape[1:], boundary_handling='zeros')
        Step = create_slab_autograd_function(op_l, dh, kernel_class=ReplayKernel)
        R = np.random.default_rng(22).standard_normal(gshape).astype(np.float32)
        u = torch.from_numpy(glob[sl].copy()).requires_grad_(True)
        (o1,) = Step.apply(u)
        (o2,) = Step.apply(o1)
        (o2 * torch.from_numpy(R[sl])).sum().backward()

        def wrapped(assigns, key, src, a):
            return evaluate(assigns, {src: np.concatenate([a[-1:], a, a[:1]], 0)}, 'zeros')[key][1:-1]
        r2 = glob.astype(np.float64)
        d = R.astype(np.float64)
        for _ in range(2):
            r2 = wrapped(op_w.forward_assignments, 'out', 'u', r2)
            d = wrapped(op_w.backward_assignments, 'diffu', 'diffout', d)
        res['autograd out'] = float(np.abs(o2.detach().numpy() - r2[sl]).max() / np.abs(r2).max())
        res['autograd grad'] = float(np.abs(u.grad.numpy() - d[sl]).max() / np.abs(d).max())
        q.put((rank, res))
    finally:
        if world > 1:
            dist.destroy_process_group()


@pytest.mark.parametrize('world', [1, 2])       # (three ranks: test_periodic_synchronisation_wraps_around)
def test_periodic_time_loops_single_steps_and_fused_pairs(world):
    """Time loops on a domain that is periodic along the decomposed axis (one rank: its own neighbour; two ranks: both
    neighbours are the same peer): ``run_steps`` synchronises the ghost planes on ONE rank too, takes fused pairs only
    where ``2 x reach`` ghost planes are stored (too few: single steps by default, an explicit request raises), and the
    ``TimeLoop`` idiom does the same; the slab autograd Function (forward and adjoint chains) wraps around too; all against
    the oracle on the wrapped global field."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_periodic_steps_worker, args=(r, world, 29840 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    for rank, res in _collect(procs, q, world):
        assert len(res) == 5 and all(v < 1e-6 for v in res.values()), (rank, res)


@pytest.mark.parametrize('fuse_steps, counts', [(False, (5, 4, 6, 7)), (None, (9, 8, 10, 11))])
def test_timeloop_cuda_graph_logic_with_recording_graphs(fuse_steps, counts):
    """The CUDA-graph path of ``TimeLoop.run`` on the CPU: a recording stand-in for ``torch.cuda.CUDAGraph``
    (tests/fake_cuda.py) captures the replayed launches with the tensors they were issued on — pointers baked in, like a
    real graph — so stale-graph mistakes (odd tail steps, external swaps, a replaced array change the buffer
    roles) show up as wrong values.  Same scenario as the GPU test ``test_timeloop_graph_follows_the_buffer_roles``, single
    steps and fused pairs, bit for bit against the eager loop."""
    from fake_cuda import FakeGraph, fake_cuda
    from pystencils_autodiff_b200.configs import make_config
    from pystencils_autodiff_b200.datahandling import SlabDataHandling
    from replay_kernels import ReplayKernel
    shape = (24, 136)
    U0 = np.random.default_rng(5).normal(size=shape).astype(np.float32)
    finals = []
    with fake_cuda(graphs=True):
        for use_graph in (True, False):
            dh = SlabDataHandling(shape, 0, 1, 0, device='cpu', backend='torch')
            dh.add_arrays('u, out', dtype=np.float32)
            kern = ReplayKernel(make_config('c2', shape=shape).forward_ast_gpu)
            dh.owned('u').copy_(torch.from_numpy(U0))
            tl = dh.create_timeloop(use_cuda_graph=use_graph, fuse_steps=fuse_steps)
            tl.add_call(kern, {})
            tl.swap('u', 'out')
            FakeGraph.replays = 0
            tl.run(counts[0])
            assert tl.fused_last_run == (fuse_steps is None)
            tl.run(counts[1])                          # roles swapped relative to the first capture
            dh.swap('u', 'out')
            dh.swap('u', 'out')
            tl.run(counts[2])
            dh.gpu_arrays['u'] = dh.owned('u').clone()  # a replaced array: same values, new pointer
            tl.run(counts[3])
            assert tl.time_steps_run == sum(counts)
            if use_graph:
                assert len(tl._graphs) >= 2 and all(g is not None for g in tl._graphs.values()) and tl.capture_error is None
                assert FakeGraph.replays >= 4           # every run() replayed at least once
            else:
                assert not tl._graphs and FakeGraph.replays == 0
            finals.append(dh.owned('u').clone())
        # a capture that fails is not fatal: the loop runs eagerly and says why
        dh = SlabDataHandling(shape, 0, 1, 0, device='cpu', backend='torch')
        dh.add_arrays('u, out', dtype=np.float32)
        dh.owned('u').copy_(torch.from_numpy(U0))

        class Failing(ReplayKernel):
            def __call__(self, **kw):
                if FakeGraph.capturing is not None:
                    raise RuntimeError('launch refused during capture')
                return ReplayKernel.__call__(self, **kw)
        tl = dh.create_timeloop(fuse_steps=fuse_steps)
        tl.add_call(Failing(make_config('c2', shape=shape).forward_ast_gpu), {})
        tl.swap('u', 'out')
        tl.run(counts[0])
        assert 'launch refused during capture' in tl.capture_error and FakeGraph.capturing is None
        eager = dh.owned('u').clone()
        dh2 = SlabDataHandling(shape, 0, 1, 0, device='cpu', backend='torch')
        dh2.add_arrays('u, out', dtype=np.float32)
        dh2.owned('u').copy_(torch.from_numpy(U0))
        dh2.run_steps(ReplayKernel(make_config('c2', shape=shape).forward_ast_gpu), counts[0], fuse=fuse_steps)
        assert torch.equal(eager, dh2.owned('u'))
    assert torch.equal(finals[0], finals[1])
    assert not torch.ones(1).is_cuda                    # the stand-ins are gone


def test_array_handler_custom_data_and_transfer_kinds():
    """The rest of the reference's data-handling vocabulary: ``array_handler`` (PyTorchArrayHandler,
    framework_integration/datahandling.py:137-174), ``add_custom_data`` with custom transfer functions
    (graph_datahandling.py:272-290) and ``DataTransferKind`` (:21-38)."""
    from pystencils_autodiff_b200.datahandling import DataTransferKind, GraphDataHandling
    dh = GraphDataHandling((6, 8), default_ghost_layers=0, device='cpu')
    h = dh.array_handler
    z, o, e, r = h.zeros((2, 3)), h.ones((2, 3), np.float64), h.empty((4,), np.float32), h.randn((3, 2))
    assert z.dtype == torch.float32 and float(z.abs().sum()) == 0 and o.dtype == torch.float64 and float(o.sum()) == 6
    assert e.shape == (4,) and r.shape == (3, 2) and r.dtype == torch.float32
    dev = h.to_gpu(np.arange(6, dtype=np.float32).reshape(2, 3))
    assert isinstance(dev, torch.Tensor) and dev.device == dh.device
    h.upload(z, np.full((2, 3), 2.0, dtype=np.float32))
    back = np.zeros((2, 3), dtype=np.float32)
    h.download(z, back)
    assert np.all(back == 2.0)
    with pytest.raises(NotImplementedError):
        h.empty((2, 3), layout='fzyx')
    # custom data with transfer functions: the data handling calls them and records markers
    log = []
    dh.add_custom_data('particles', lambda: {'host': [1, 2, 3]}, lambda: {'device': []},
                       lambda gpu, cpu: (gpu.__setitem__('device', list(cpu['host'])), log.append('up')),
                       lambda gpu, cpu: (cpu.__setitem__('host', list(gpu['device'])), log.append('down')))
    dh.to_gpu('particles')
    dh.custom_data_gpu['particles']['device'].append(4)
    dh.to_cpu('particles')
    assert dh.custom_data_cpu['particles']['host'] == [1, 2, 3, 4] and log == ['up', 'down']
    assert [c[0] for c in dh.call_queue] == ['CustomData', 'CustomTransfer', 'CustomTransfer']
    with pytest.raises(ValueError, match='both transfer functions'):
        dh.add_custom_data('half', dict, dict, cpu_to_gpu_transfer_func=lambda g, c: None)
    with pytest.raises(ValueError, match='already been added'):
        dh.add_custom_data('particles', dict)
    # plain arrays keep their recorded transfers, named after the enum's members
    dh.add_arrays('u')
    dh.to_cpu('u')
    dh.to_gpu('u')
    kinds = [DataTransferKind(c[2]) for c in dh.call_queue if c[0] == 'DataTransfer']
    assert kinds == [DataTransferKind.DEVICE_TO_HOST, DataTransferKind.HOST_TO_DEVICE]
    assert all(k.is_transfer() and not k.is_alloc() for k in kinds) and DataTransferKind.DEVICE_ALLOC.is_alloc()
    assert DataTransferKind.DEVICE_SWAP.is_transfer() and not DataTransferKind.HOST_GATHER.is_transfer()
    from pystencils_autodiff_b200.computationgraph import ComputationGraph
    assert len(ComputationGraph(dh).computation_nodes) == 2          # the custom-data markers carry no array dependencies
