"""GPU: operator-API behaviour (marshalling rules of backends/_torch_native.py) and full-size property checks."""
import numpy as np
import pytest
import sympy as sp

import pystencils_autodiff_b200 as ps
from oracle import evaluate, forward_backward
from pystencils_autodiff_b200.configs import make_config

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_scalar_parameter_through_class_kwargs():
    """tests/backends/test_torch_native_compilation.py:156-211: z = x*log(a*x*y), a = 5, rand(20,40), all cells."""
    import torch
    z, y, x = ps.fields("z, y, x: [20,40]")
    a = sp.Symbol('a')
    op = ps.AutoDiffOp(ps.AssignmentCollection({z[0, 0]: x[0, 0] * sp.log(a * x[0, 0] * y[0, 0])}), op_name='scal')
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    fn.class_kwargs['a'] = 5.0
    rng = np.random.default_rng(0)
    X, Y, G = rng.random((20, 40)) + 0.1, rng.random((20, 40)) + 0.1, rng.normal(size=(20, 40))
    xt, yt = _t(X).requires_grad_(True), _t(Y).requires_grad_(True)
    (out,) = fn.apply(xt, yt)
    out.backward(_t(G))
    ref_o, ref_d = forward_backward(op, dict(x=X, y=Y), dict(z=G), scalars=dict(a=5.0))
    assert np.abs(out.detach().cpu().numpy() - ref_o['z']).max() <= 1e-12 * np.abs(ref_o['z']).max()
    assert np.abs(xt.grad.cpu().numpy() - ref_d['diffx']).max() <= 1e-12 * np.abs(ref_d['diffx']).max()
    assert np.abs(yt.grad.cpu().numpy() - ref_d['diffy']).max() <= 1e-12 * np.abs(ref_d['diffy']).max()
    assert [p.symbol.name for p in fn.forward_parameters] == ['x', 'y']
    assert fn.call(x=xt, y=yt).shape == (20, 40)
    assert 'psad_scal_forward_gpu' in fn.code
    with pytest.raises(TypeError):
        fn.apply(xt.float(), yt)            # dtype mismatch is an error, not a reinterpretation


def test_constant_fields_get_no_gradient_and_outputs_are_a_tuple():
    import torch
    x, y, z = ps.fields('x, y, z: float64[12,16]')
    fa = ps.AssignmentCollection({z.center: x[1, 0] * y[0, 0] + x[0, -1]})
    op = ps.AutoDiffOp(fa, boundary_handling='zeros', constant_fields=[y])
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    rng = np.random.default_rng(1)
    X, Y, G = rng.normal(size=(12, 16)), rng.normal(size=(12, 16)), rng.normal(size=(12, 16))
    xt, yt = _t(X).requires_grad_(True), _t(Y).requires_grad_(True)
    outs = fn.apply(xt, yt)
    assert isinstance(outs, tuple) and len(outs) == 1
    outs[0].backward(_t(G))
    assert yt.grad is None
    _, ref = forward_backward(op, dict(x=X, y=Y), dict(z=G))
    np.testing.assert_allclose(xt.grad.cpu().numpy(), ref['diffx'], rtol=1e-12, atol=1e-13)


def test_time_constant_field_accumulate_form():
    """``diff_f.center += ...`` (_autodiff.py:110-113): the kernel reads its own (zero-initialised) output."""
    import torch
    u, out = ps.fields('u, out: float64[10,32]')
    fa = [ps.Assignment(out.center, 0.5 * u[0, 1] - 2 * u[1, 0] + u[0, 0])]
    op = ps.AutoDiffOp(fa, boundary_handling='zeros', time_constant_fields=[u])
    assert 'diffu_C' in str(op.backward_assignments)
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    ut = torch.randn(10, 32, dtype=torch.float64, device='cuda', requires_grad=True)
    assert torch.autograd.gradcheck(fn.apply, (ut,), atol=1e-4)


def test_double_compute_type_for_float_fields():
    """pystencils' default evaluates float32 fields in double; ``data_type='double'`` reproduces that."""
    shape = (24, 128)
    rng = np.random.default_rng(2)
    U = rng.normal(size=shape).astype(np.float32)
    res = {}
    for dt in (None, 'double'):
        op = make_config('c2', shape=shape, **({'data_type': dt} if dt else {}))
        fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
        (o,) = fn.apply(_t(U))
        res[dt] = o.cpu().numpy()
        assert ('typedef double CT' in fn.code) == (dt == 'double')
    ref = evaluate(make_config('c2', shape=shape).forward_assignments, dict(u=U), 'zeros')['out']
    # double arithmetic, one rounding on store: at most 1 ulp of float32 away from the float64 oracle (FMA vs mul+add)
    assert np.abs(res['double'] - ref).max() <= 1.2e-7 * np.abs(ref).max()
    assert np.abs(res[None] - ref).max() <= 1e-6 * np.abs(ref).max()


@pytest.mark.parametrize('shape', [(1, 1, 4), (1, 3, 4), (2, 2, 8), (3, 1, 128), (5, 33, 132)])
@pytest.mark.parametrize('bh', [None, 'zeros'])
def test_degenerate_shapes(shape, bh):
    """Arrays thinner than the stencil / the tile, single planes, single rows."""
    import torch
    op = make_config('c3', shape=shape, boundary_handling=bh)
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    rng = np.random.default_rng(3)
    U, G = rng.normal(size=shape).astype(np.float32), rng.normal(size=shape).astype(np.float32)
    ut = _t(U).requires_grad_(True)
    (o,) = fn.apply(ut)
    o.backward(_t(G))
    ref_o, ref_d = forward_backward(op, dict(u=U), dict(out=G))
    scale = max(1e-30, np.abs(ref_o['out']).max())
    assert np.abs(o.detach().cpu().numpy() - ref_o['out']).max() <= 1e-6 * max(scale, 1.0)
    assert np.abs(ut.grad.cpu().numpy() - ref_d['diffu']).max() <= 1e-6 * max(1.0, np.abs(ref_d['diffu']).max())


def test_non_contiguous_inputs_and_non_default_stream():
    import torch
    op = make_config('c2', shape=(64, 128))
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    base = torch.randn(128, 64, device='cuda')
    u = base.t()                                      # non-contiguous view
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        (o1,) = fn.apply(u)
    s.synchronize()
    (o2,) = fn.apply(u.contiguous())
    assert torch.equal(o1, o2)


@pytest.mark.parametrize('name', ['c2', 'c3', 'c4'])
def test_full_size_adjoint_identity_and_linearity(name):
    """At the BASELINE sizes the oracle is too slow; use size-independent properties of linear stencils in 'zeros'
    mode: <A x, y> == <x, A^T y> (the adjoint kernel is the exact transpose) and A(a x1 + x2) == a A x1 + A x2."""
    import torch
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    op = make_config(name)
    shape = tuple(int(s) for s in op.forward_input_fields[0].shape)
    dt = torch.float32 if name != 'c4' else torch.float64
    fk, bk = CompiledKernel(op.forward_ast_gpu), CompiledKernel(op.backward_ast_gpu)
    g = torch.Generator(device='cuda')
    g.manual_seed(11)
    x = torch.randn(shape, dtype=dt, device='cuda', generator=g)
    y = torch.randn(shape, dtype=dt, device='cuda', generator=g)
    Ax, Aty = torch.empty_like(x), torch.empty_like(x)
    fk(u=x, out=Ax)
    bk(diffout=y, diffu=Aty)
    assert fk.last_variant == 'march' and bk.last_variant == 'march'
    lhs = torch.dot(Ax.double().flatten(), y.double().flatten()).item()
    rhs = torch.dot(x.double().flatten(), Aty.double().flatten()).item()
    tol = 1e-5 if name != 'c4' else 1e-12
    assert abs(lhs - rhs) <= tol * (abs(lhs) + abs(rhs) + np.sqrt(x.numel()))
    # linearity
    x2 = torch.randn(shape, dtype=dt, device='cuda', generator=g)
    Ax2, Acomb = torch.empty_like(x), torch.empty_like(x)
    fk(u=x2, out=Ax2)
    fk(u=0.5 * x + x2, out=Acomb)
    err = (Acomb - (0.5 * Ax + Ax2)).abs().max().item()
    assert err <= (2e-6 if name != 'c4' else 1e-14) * 4
    # checksum of checksums: sum(A x) == <x, A^T 1> (column sums: 1 in the interior, less on the zero boundary)
    ones, At1 = torch.ones_like(x), torch.empty_like(x)
    bk(diffout=ones, diffu=At1)
    lhs, rhs = Ax.double().sum().item(), torch.dot(x.double().flatten(), At1.double().flatten()).item()
    assert abs(lhs - rhs) <= tol * (abs(lhs) + abs(rhs) + np.sqrt(x.numel()))


def test_tv_full_size_is_finite_and_matches_generic_on_a_window():
    import torch
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    op = make_config('c5', shape=(2, 4096, 4096))
    fk = CompiledKernel(op.forward_ast_gpu)
    g = torch.Generator(device='cuda')
    g.manual_seed(5)
    u = torch.rand((2, 4096, 4096), device='cuda', generator=g)
    f = torch.rand((2, 4096, 4096), device='cuda', generator=g)
    o1, o2 = torch.empty_like(u), torch.empty_like(u)
    fk(u=u, f=f, g=o1)
    fk(u=u, f=f, g=o2, _variant='generic')
    assert torch.isfinite(o1).all()
    assert (o1 - o2).abs().max().item() <= 5e-5 * o2.abs().max().item()


@pytest.mark.parametrize('name,shape,T', [('c4', (10, 14, 64), 4), ('c3', (12, 18, 128), 3)])
def test_multi_timestep_unrolled_autograd(name, shape, T):
    """BASELINE config 4: T unrolled ``op.apply`` calls, loss = sum(out_T * r); gradient w.r.t. the initial field
    against the oracle chain (forward T times, adjoint T times in reverse)."""
    import torch
    op = make_config(name, shape=shape, boundary_handling='zeros')
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    rng = np.random.default_rng(9)
    dt = op.forward_input_fields[0].dtype.numpy_dtype
    U0, Rw = rng.normal(size=shape).astype(dt), rng.normal(size=shape).astype(dt)
    u = _t(U0).requires_grad_(True)
    x = u
    for _ in range(T):
        (x,) = fn.apply(x)
    loss = (x * _t(Rw)).sum()
    loss.backward()
    ref = U0.astype(np.float64)
    for _ in range(T):
        ref = evaluate(op.forward_assignments, dict(u=ref.astype(dt)), 'zeros')['out'].astype(np.float64)
    g = Rw.astype(np.float64)
    for _ in range(T):
        g = evaluate(op.backward_assignments, dict(diffout=g.astype(dt)), 'zeros')['diffu'].astype(np.float64)
    tol = 1e-12 if dt == np.float64 else 2e-6
    assert np.abs(x.detach().cpu().numpy() - ref).max() <= tol * np.abs(ref).max()
    assert np.abs(u.grad.cpu().numpy() - g).max() <= tol * np.abs(g).max()
    # linear stencil: nothing is saved for backward, intermediates can be freed as the chain advances
    assert fn.backward_kernel.ir.input_fields[0].name == 'diffout' and len(fn.backward_kernel.ir.input_fields) == 1


def test_cuda_graph_capture_of_forward_and_adjoint():
    """Launches go through cuLaunchKernel on torch's current stream, so a step can be captured in a CUDA graph and
    replayed (launch-bound small fields)."""
    import torch
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    op = make_config('c2', shape=(256, 256))
    fk, bk = CompiledKernel(op.forward_ast_gpu), CompiledKernel(op.backward_ast_gpu)
    u = torch.randn(256, 256, device='cuda')
    out, du = torch.empty_like(u), torch.empty_like(u)
    fk(u=u, out=out)
    bk(diffout=out, diffu=du)
    torch.cuda.synchronize()
    ref = du.clone()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            fk(u=u, out=out)
            bk(diffout=out, diffu=du)
    du.zero_()
    gr.replay()
    torch.cuda.synchronize()
    assert torch.equal(du, ref)


def test_timeloop_with_swap_and_cuda_graph_replay():
    """TimeLoop (graph_datahandling.py:152-194 API) over SlabDataHandling: kernel + swap per step, replayed from a
    CUDA graph, equals the same steps issued eagerly and the oracle chain."""
    import torch
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    from pystencils_autodiff_b200.datahandling import SlabDataHandling
    shape, T = (64, 128), 11
    op = make_config('c2', shape=shape)
    kern = CompiledKernel(op.forward_ast_gpu)
    rng = np.random.default_rng(4)
    U0 = rng.normal(size=shape).astype(np.float32)
    results = []
    for use_graph in (True, False):
        dh = SlabDataHandling(shape, 0, 1, 0, device='cuda:0')
        dh.add_arrays('u, out', dtype=np.float32)
        dh.owned('u').copy_(_t(U0))
        tl = dh.create_timeloop(use_cuda_graph=use_graph)
        tl.add_call(kern, {})
        tl.add_single_step_function(lambda d=dh: d.swap('u', 'out'))
        tl.run(T)
        torch.cuda.synchronize()
        assert tl.time_steps_run == T
        results.append(dh.owned('u').clone())
        # graph_datahandling.py:181-190: the run is ONE queue entry holding what a step consists of
        kind, steps, recorded = dh.call_queue[-1]
        assert kind == 'TimeloopRun' and steps == T and [c[0] for c in recorded] == ['KernelCall', 'Swap']
        assert not any(c[0] == 'KernelCall' for c in dh.call_queue)
    assert torch.equal(results[0], results[1])
    ref = U0
    for _ in range(T):
        ref = evaluate(op.forward_assignments, dict(u=ref), 'zeros')['out']
    assert np.abs(results[0].cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize('fuse_steps, counts', [(False, (5, 4, 6, 7)), (None, (9, 8, 10, 11))])
def test_timeloop_graph_follows_the_buffer_roles(fuse_steps, counts):
    """a captured graph bakes buffer pointers in.  ``run(5)`` leaves ``u`` / ``out`` swapped after its odd
    eager tail step, an external ``swap`` or a replaced array changes the roles too — each must re-capture (or reuse the
    graph captured for exactly those roles), never replay a stale one.  Second case: the same with the loop issuing fused
    pairs of steps (the graph then holds two pairs = four time steps)."""
    import torch
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    from pystencils_autodiff_b200.datahandling import SlabDataHandling
    shape = (64, 128)
    op = make_config('c2', shape=shape)
    kern = CompiledKernel(op.forward_ast_gpu)
    U0 = np.random.default_rng(5).normal(size=shape).astype(np.float32)
    finals = []
    for use_graph in (True, False):
        dh = SlabDataHandling(shape, 0, 1, 0, device='cuda:0')
        dh.add_arrays('u, out', dtype=np.float32)
        dh.owned('u').copy_(_t(U0))
        tl = dh.create_timeloop(use_cuda_graph=use_graph, fuse_steps=fuse_steps)
        tl.add_call(kern, {})
        tl.swap('u', 'out')
        tl.run(counts[0])
        assert tl.fused_last_run == (fuse_steps is None)
        tl.run(counts[1])                          # roles swapped relative to the first capture
        dh.swap('u', 'out')
        dh.swap('u', 'out')
        tl.run(counts[2])
        fresh = dh.owned('u').clone()              # a replaced array: same values, new pointer
        dh.gpu_arrays['u'] = fresh
        tl.run(counts[3])
        torch.cuda.synchronize()
        assert tl.time_steps_run == sum(counts)
        if use_graph and torch.cuda.is_available():          # (the CPU dry run of this body has no CUDA graphs)
            assert len(tl._graphs) >= 2 and all(g is not None for g in tl._graphs.values())
        finals.append(dh.owned('u').clone())
    assert torch.equal(finals[0], finals[1])
    ref = U0
    for _ in range(sum(counts)):
        ref = evaluate(op.forward_assignments, dict(u=ref), 'zeros')['out']
    assert np.abs(finals[0].cpu().numpy() - ref).max() <= 4e-6 * np.abs(ref).max()


@pytest.mark.parametrize('name, shape, g', [('c3', (20, 30, 128), 2), ('c3', (20, 30, 128), 0), ('c2', (64, 128), 0)])
def test_timeloop_runs_fused_pairs_of_steps(name, shape, g):
    """f-1 through the reference's API: ``add_call(kernel)`` + ``swap(in, out)`` runs as ``out = S(S(u))`` launches (two time
    steps each, CUDA-graph replayed) where ``run_steps`` fuses by default — bit-identical to ``run_steps``, half the launches,
    the oracle chain within tolerance; ``fuse_steps=False`` issues single steps."""
    import torch
    from pystencils_autodiff_b200 import runtime
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    from pystencils_autodiff_b200.datahandling import SlabDataHandling
    T = 13
    op = make_config(name, shape=shape)
    U0 = np.random.default_rng(6).normal(size=shape).astype(np.float32)
    finals, launches = {}, {}
    for mode in (None, False, 'run_steps'):
        dh = SlabDataHandling(shape, 0, 1, g, device='cuda:0')
        dh.add_arrays('u, out', dtype=np.float32)
        kern = CompiledKernel(make_config(name, shape=dh.dec.local_shape).forward_ast_gpu)
        dh.owned('u').copy_(_t(U0))
        n0 = runtime.launch_count()
        if mode == 'run_steps':
            dh.run_steps(kern, T)
        else:
            tl = dh.create_timeloop(use_cuda_graph=False, fuse_steps=mode)
            tl.add_call(kern, {})
            tl.swap('u', 'out')
            tl.run(T)
            assert tl.fused_last_run == (mode is None)
        torch.cuda.synchronize()
        launches[mode] = runtime.launch_count() - n0
        finals[mode] = dh.owned('u').clone()
        if mode is None:                       # the same loop replayed from a CUDA graph of two pairs
            dh.owned('u').copy_(_t(U0))
            tg = dh.create_timeloop(use_cuda_graph=True)
            tg.add_call(kern, {})
            tg.swap('u', 'out')
            tg.run(T)
            torch.cuda.synchronize()
            assert tg.fused_last_run
            if torch.cuda.is_available():          # (the CPU dry run of this body has no CUDA graphs)
                assert len(tg._graphs) == 1 and all(v is not None for v in tg._graphs.values())
            assert torch.equal(dh.owned('u'), finals[None])
    if torch.cuda.is_available():              # (replayed launches of the CPU dry run are not counted by the runtime)
        assert launches[None] == launches['run_steps'] == T // 2 + 1 and launches[False] == T
    assert torch.equal(finals[None], finals['run_steps'])
    ref = U0.astype(np.float64)
    for _ in range(T):
        ref = evaluate(op.forward_assignments, dict(u=ref), 'zeros')['out']
    for mode in (None, False):
        assert np.abs(finals[mode].cpu().numpy() - ref).max() <= 4e-6 * np.abs(ref).max()


def test_tensor_field_front_door():
    import torch
    from pystencils_autodiff_b200.field_tensor_conversion import (coerce_to_field, create_field_from_array_like,
                                                                   is_array_like, torch_tensor_from_field)
    a, b = torch.zeros((20, 10)), torch.zeros((6, 7), dtype=torch.float64)
    x, y = ps.fields(x=a, y=b)          # tests/backends/test_torch_native_compilation.py:247-256
    assert x.shape == (20, 10) and y.dtype.numpy_dtype == np.float64
    c = torch.zeros((20, 10)).cuda()
    z = ps.fields(z=c)
    assert z.shape == (20, 10) and is_array_like(c) and not is_array_like(z)
    assert coerce_to_field('w', c).name == 'w' and create_field_from_array_like('q', b).strides == (7, 1)
    t = torch_tensor_from_field(z, init_val=2.0, cuda=True)
    assert t.is_cuda and t.shape == (20, 10) and float(t[0, 0]) == 2.0


@pytest.mark.parametrize('name,shape,lo', [('c1', (24, 32), 0.5), ('c2', (48, 128), -1), ('c5', (2, 32, 128), 0.0),
                                           ('c3', (10, 30, 128), -1), ('c1', (20, 30), 0.5)])
@pytest.mark.parametrize('bh', [None, 'zeros'])
def test_fused_forward_adjoint_equals_two_kernels(name, shape, lo, bh):
    """One launch over the union of forward and adjoint assignments == the two separate kernels."""
    import torch
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    op = make_config(name, shape=shape, boundary_handling=bh)
    fused = op.fused_kernel_gpu
    assert fused.ir.bytes_per_cell() <= op.forward_ast_gpu.bytes_per_cell() + op.backward_ast_gpu.bytes_per_cell()
    g = torch.Generator(device='cuda')
    g.manual_seed(3)
    tens = {}
    for f in fused.ir.input_fields:
        tens[f.name] = torch.rand(shape, dtype=getattr(torch, f.dtype.numpy_dtype.name), device='cuda', generator=g) + lo + 0.01
    a = dict(tens)
    for f in fused.ir.output_fields:
        a[f.name] = torch.full(shape, float('nan'), dtype=tens[fused.ir.input_fields[0].name].dtype, device='cuda')
    fused(**{f.name: a[f.name] for f in fused.fields})
    b = dict(tens)
    for k in (CompiledKernel(op.forward_ast_gpu), CompiledKernel(op.backward_ast_gpu)):
        for f in k.ir.output_fields:
            b[f.name] = torch.full(shape, float('nan'), dtype=a[f.name].dtype, device='cuda')
        k(**{f.name: b[f.name] for f in k.fields})
    for f in fused.ir.output_fields:
        # same per-cell expressions; the summation order may differ (the fused kernel has more accesses and may
        # switch to arrival-time plane sums), so allow rounding-level differences
        scale = b[f.name].abs().max().item()
        assert (a[f.name] - b[f.name]).abs().max().item() <= 2e-6 * max(scale, 1e-30), f.name


def test_vector_output_through_the_function_uses_soa_and_fast_path():
    """Curl (vector output): the Function allocates the output structure-of-arrays, both kernels take the march path,
    gradcheck-style comparison against the oracle."""
    import torch
    shape = (24, 128)
    u = ps.Field.create_fixed_size('curl_input', shape, index_dimensions=0, dtype=np.float64)
    c = ps.Field.create_fixed_size('curl', shape + (2,), index_dimensions=1, dtype=np.float64)
    disc = ps.fd.Discretization2ndOrder(dx=1)
    fa = ps.AssignmentCollection([ps.Assignment(c.center(0), disc(ps.fd.Diff(u, 0))),
                                  ps.Assignment(c.center(1), disc(ps.fd.Diff(u, 1)) + u.center ** 2)], [])
    op = ps.AutoDiffOp(fa, boundary_handling='zeros')
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    rng = np.random.default_rng(6)
    U, G = rng.normal(size=shape), rng.normal(size=shape + (2,))
    ut = _t(U).requires_grad_(True)
    (curl,) = fn.apply(ut)
    assert tuple(curl.shape) == shape + (2,) and curl.stride(1) == 1
    curl.backward(_t(G))
    assert fn.forward_kernel.last_variant == 'march' and fn.backward_kernel.last_variant == 'march'
    ref_o, ref_d = forward_backward(op, dict(curl_input=U), dict(curl=G))
    np.testing.assert_allclose(curl.detach().cpu().numpy(), ref_o['curl'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ut.grad.cpu().numpy(), ref_d['diffcurl_input'], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('with_offsets', (False, True))
def test_tfmad_gradient_check_torch_native_reference_case(with_offsets):
    """/root/reference/tests/test_tfmad.py:186-231 verbatim (float64[5,7], 'zeros', gradcheck atol=1e-4), plus the
    same check on random non-zero inputs (the reference only ever checks all-zero tensors)."""
    import torch
    a, b, out = ps.fields("a, b, out: float64[5,7]")
    if with_offsets:
        cont = 2 * ps.fd.Diff(a, 0) - 1.5 * ps.fd.Diff(a, 1) - ps.fd.Diff(b, 0) + 3 * ps.fd.Diff(b, 1)
        assignment = ps.Assignment(out.center(), ps.fd.Discretization2ndOrder(dx=1)(cont) + 1.2 * a.center())
    else:
        assignment = ps.Assignment(out.center(), 1.2 * a.center + 0.1 * b.center)
    auto_diff = ps.AutoDiffOp(ps.AssignmentCollection([assignment], []), boundary_handling='zeros',
                              diff_mode='transposed-forward')
    function = auto_diff.create_tensorflow_op(use_cuda=True, backend='torch_native')
    for maker in (torch.zeros, torch.randn):
        a_tensor = maker(*a.shape, dtype=torch.float64, device='cuda').requires_grad_(True)
        b_tensor = maker(*b.shape, dtype=torch.float64, device='cuda').requires_grad_(True)
        d = {a: a_tensor, b: b_tensor}
        assert torch.autograd.gradcheck(function.apply, tuple(d[f] for f in auto_diff.forward_input_fields),
                                        atol=1e-4, raise_exception=True)


def test_tfmad_gradient_check_two_outputs_reference_case():
    """/root/reference/tests/test_tfmad.py:234-285: three outputs incl. exp(b[-1,0]), float64[21,13], 'zeros'.
    With the reference's all-zero inputs gradcheck passes; on random inputs the adjoint is the reference's
    (un-shifted coefficient) and is compared with the oracle instead."""
    import torch
    a, b, out1, out2, out3 = ps.fields("a, b, out1, out2, out3: float64[21,13]")
    ac = ps.AssignmentCollection({out1.center: a.center + b.center, out2.center: a.center - b.center,
                                  out3.center: sp.exp(b[-1, 0])})
    auto_diff = ps.AutoDiffOp(ac, boundary_handling='zeros', diff_mode='transposed-forward')
    function = auto_diff.create_tensorflow_op(use_cuda=True, backend='torch_native')
    a_tensor = torch.zeros(*a.shape, dtype=torch.float64, device='cuda', requires_grad=True)
    b_tensor = torch.zeros(*b.shape, dtype=torch.float64, device='cuda', requires_grad=True)
    assert torch.autograd.gradcheck(function.apply, (a_tensor, b_tensor), atol=1e-4, raise_exception=True)
    rng = np.random.default_rng(12)
    A, B = rng.normal(size=(21, 13)), rng.normal(size=(21, 13))
    G = {n: rng.normal(size=(21, 13)) for n in ('out1', 'out2', 'out3')}
    at, bt = _t(A).requires_grad_(True), _t(B).requires_grad_(True)
    outs = function.apply(at, bt)
    assert len(outs) == 3
    torch.autograd.backward(outs, [_t(G[f.name]) for f in auto_diff.forward_output_fields])
    ref_o, ref_d = forward_backward(auto_diff, dict(a=A, b=B), G)
    for f, o in zip(auto_diff.forward_output_fields, outs):
        np.testing.assert_allclose(o.detach().cpu().numpy(), ref_o[f.name], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(at.grad.cpu().numpy(), ref_d['diffa'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(bt.grad.cpu().numpy(), ref_d['diffb'], rtol=1e-12, atol=1e-12)


def test_host_streamed_op_with_scalar_and_accumulate_form():
    """HostStreamedOp passes scalar parameters through and zero-initialises ``+=``-form outputs per chunk."""
    import torch
    from pystencils_autodiff_b200.datahandling import HostStreamedOp
    shape = (21, 16, 64)
    u, out = ps.fields('u, out: float64[%d,%d,%d]' % shape)
    al = sp.Symbol('alpha')
    fa = [ps.Assignment(out.center, u[0, 0, 0] + al * (u[1, 0, 0] + u[-1, 0, 0] + u[0, 0, 1] - 3 * u[0, 0, 0]))]
    op = ps.AutoDiffOp(fa, boundary_handling='zeros', time_constant_fields=[u])
    st = HostStreamedOp(op, shape, 'cuda:0', chunk_planes=4)
    rng = np.random.default_rng(21)
    U, G = rng.normal(size=shape), rng.normal(size=shape)
    host = {'u': torch.from_numpy(U).pin_memory(), 'diffout': torch.from_numpy(G).pin_memory(),
            'out': torch.empty(shape, dtype=torch.float64).pin_memory(),
            'diffu': torch.empty(shape, dtype=torch.float64).pin_memory()}
    assert sorted(st.input_names) == ['diffout', 'u'] and sorted(st.output_names) == ['diffu', 'out']
    with pytest.raises(TypeError):
        st({n: host[n] for n in st.input_names}, {n: host[n] for n in st.output_names})
    st({n: host[n] for n in st.input_names}, {n: host[n] for n in st.output_names}, alpha=0.3)
    torch.cuda.synchronize()
    ref_o, ref_d = forward_backward(op, dict(u=U), dict(out=G), scalars=dict(alpha=0.3))
    np.testing.assert_allclose(host['out'].numpy(), ref_o['out'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(host['diffu'].numpy(), ref_d['diffu'], rtol=1e-12, atol=1e-12)


def test_gradients_are_matched_to_outputs_by_name_not_position():
    """An output whose value does not depend on any input has no ``diff<out>`` field in the adjoint kernel; the
    gradients of the remaining outputs must still reach the right fields (the reference binds them by position)."""
    import torch
    x, a_out, b_out = ps.fields('x, a_out, b_out: float64[8,16]')
    ac = ps.AssignmentCollection({a_out.center: sp.Float(2.5) + 0 * x.center, b_out.center: 3 * x[0, 1] + x.center})
    op = ps.AutoDiffOp(ac, boundary_handling='zeros')
    assert [f.name for f in op.backward_input_fields] == ['diffb_out']
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    rng = np.random.default_rng(2)
    X, GA, GB = rng.normal(size=(8, 16)), rng.normal(size=(8, 16)), rng.normal(size=(8, 16))
    xt = _t(X).requires_grad_(True)
    oa, ob = fn.apply(xt)
    torch.autograd.backward((oa, ob), (_t(GA), _t(GB)))
    _, ref = forward_backward(op, dict(x=X), dict(a_out=GA, b_out=GB))
    np.testing.assert_allclose(xt.grad.cpu().numpy(), ref['diffx'], rtol=1e-13, atol=1e-13)


def test_repeated_launches_hit_the_launch_cache_and_stay_correct():
    """psad_kernel_launch remembers the parameter block + encoded tensor maps per (buffers, shapes, strides, range) and
    CompiledKernel.__call__ the validated argument pack: repeats must hit, give the same bits, follow a changed scalar,
    and new buffers / views must miss (never reuse another buffer's tensor map)."""
    import torch
    from pystencils_autodiff_b200 import runtime
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    import pystencils_autodiff_b200 as ps
    u, out = ps.fields('u, out: float32[48,128]')
    a = sp.Symbol('a')
    op = ps.AutoDiffOp([ps.Assignment(out.center, a * (u[0, 1] + u[0, -1] + u[1, 0] + u[-1, 0]) - u.center)],
                       boundary_handling='zeros')
    k = CompiledKernel(op.forward_ast_gpu)
    rng = np.random.default_rng(0)
    U = [rng.normal(size=(48, 128)).astype(np.float32) for _ in range(3)]
    tu = [_t(x) for x in U]
    to = [torch.empty_like(tu[0]) for _ in range(3)]
    h0, m0 = runtime.launch_cache_stats()
    for rep in range(3):
        for i in range(3):
            k(u=tu[i], out=to[i], a=0.5 + rep)
    torch.cuda.synchronize()
    h1, m1 = runtime.launch_cache_stats()
    assert m1 - m0 == 3 and h1 - h0 == 6                  # three buffer pairs, each built once
    for i in range(3):
        ref = evaluate(op.forward_assignments, dict(u=U[i]), 'zeros', scalars={'a': 2.5})['out']
        assert np.abs(to[i].cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()
    # a view with the same pointer but another shape is another launch
    k2 = CompiledKernel(ps.AutoDiffOp([ps.Assignment(ps.fields('o2: float32[24,128]').center,
                                                     ps.fields('u2: float32[24,128]')[1, 0])],
                                      boundary_handling='zeros').forward_ast_gpu)
    o2 = torch.empty((24, 128), device='cuda')
    k2(u2=tu[0][:24], o2=o2)
    k2(u2=tu[0][24:], o2=o2)
    torch.cuda.synchronize()
    assert torch.equal(o2[:-1], tu[0][25:]) and not o2[-1].any()


def test_iteration_range_outside_the_array_is_rejected():
    """build_args validated only the write range.  The generic kernel with interior iteration reads unguarded:
    an iteration range closer than the ghost width to the array edge must be an error, not an out-of-bounds read."""
    import torch
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    op = make_config('c2', shape=(20, 30), boundary_handling=None)           # 120-byte rows: generic kernel
    k = CompiledKernel(op.forward_ast_gpu)
    u = torch.zeros((20, 30), device='cuda')
    out = torch.empty_like(u)
    with pytest.raises(RuntimeError, match='iteration range'):
        k(u=u, out=out, _range=dict(iter_lo=[0, 1], iter_hi=[20, 29], write_lo=[0, 0], write_hi=[20, 30]))
    with pytest.raises(RuntimeError, match='iteration range'):
        k(u=u, out=out, _range=dict(iter_lo=[1, 1], iter_hi=[19, 31], write_lo=[0, 0], write_hi=[20, 30]))
    k(u=u, out=out, _range=dict(iter_lo=[1, 1], iter_hi=[19, 29], write_lo=[0, 0], write_hi=[20, 30]))
    torch.cuda.synchronize()


def test_backward_that_reads_nothing_still_allocates_its_outputs():
    """a user-supplied adjoint with constant right-hand sides reads neither an upstream gradient nor a saved
    tensor; the shape template then comes from the forward call (used to raise a bare StopIteration)."""
    import torch
    from pystencils_autodiff_b200._adjoint_field import AdjointField
    u, out = ps.fields('u, out: float32[2D]')
    op = ps.AutoDiffOp([ps.Assignment(out.center, 2 * u[0, 1])],
                       backward_assignments=[ps.Assignment(AdjointField(u).center, sp.Float(3.0))], boundary_handling='zeros')
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    x = torch.ones((8, 16), device='cuda', requires_grad=True)
    (y,) = fn.apply(x)
    y.sum().backward()
    assert x.grad.shape == (8, 16) and float(x.grad.min()) == 3.0 and float(x.grad.max()) == 3.0


@pytest.mark.parametrize('name', ['c2', 'c3', 'c4', 'c5'])
def test_full_size_outputs_match_the_c_oracle_on_sampled_blocks(name):
    """At the BASELINE.json sizes: blocks of planes (first, middle, last; whole images for C5) of what the forward and the
    adjoint kernel produced, recomputed by the C restatement of the pystencils CPU loop nest (oracle/cgen.py, 'strict': double
    precision, no contraction) from the same inputs plus their halo planes.  North-star tolerances: 1e-6 (fp32), 1e-12
    (fp64), norm-wise — with ONE stated exception: the adjoint of the TV gradient in fp32 arithmetic, 1e-5 (a few dozen of its
    268 M cells have vanishing gradients, coefficients ~1/|grad u|^3 and cancelling terms of +-1e3: bench.py ``TOL_OVERRIDE``);
    in the reference's own arithmetic (``data_type='double'``) it meets 1e-6 like everything else, checked below.
    (The same checks run inside bench.py's line as ``parity`` / ``double_arithmetic``.)"""
    import types
    import torch
    import bench
    from pystencils_autodiff_b200.configs import CONFIG_SHAPES
    from pystencils_autodiff_b200.datahandling import SlabStencilOp
    shape = tuple(CONFIG_SHAPES[name]['shape'])
    op = make_config(name, shape=shape, boundary_handling='zeros')
    slab = SlabStencilOp(op, local_shape=shape, rank=0, world_size=1, device=torch.device('cuda', 0))
    g = torch.Generator(device='cuda')
    g.manual_seed(77)
    slab.randomize(g)
    res = bench.oracle_parity(types.SimpleNamespace(torch=torch), name, op, slab)
    assert slab.fwd.last_variant == 'march' and slab.bwd.last_variant == 'march'
    assert res['ok'], res
    for tag, err in res['per_kernel'].items():
        assert err <= (1e-5 if (name, tag) == ('c5', 'adjoint') else res['tolerance']), (tag, res)
    assert res['max_rel_err'] > 0 or name == 'c4'          # an fp32 kernel that matches a double oracle exactly compared nothing
    assert len(res['blocks_dim0']) == 6
    if name == 'c5':
        dbl = bench.double_arithmetic(types.SimpleNamespace(torch=torch), name, slab, shape)
        assert dbl['parity']['ok'] and dbl['parity']['max_rel_err'] <= 1e-6, dbl


@pytest.mark.parametrize('bh', ['zeros', None])
def test_tv_gradient_in_double_arithmetic_meets_the_fp32_tolerance(bh):
    """C5 with ``data_type='double'`` (pystencils' default promotion for fp32 fields) on the GPU, forward and
    adjoint, both boundary modes, at the north-star tolerance."""
    import torch
    shape = (3, 72, 256)
    op = make_config('c5', shape=shape, boundary_handling=bh, data_type='double')
    rng = np.random.default_rng(21)
    ins = {f.name: rng.uniform(0, 1, shape).astype(np.float32) for f in op.forward_input_fields}
    grads = {f.name: rng.standard_normal(shape).astype(np.float32) for f in op.forward_output_fields}
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    xs = [_t(ins[f.name]).requires_grad_(True) for f in op.forward_input_fields]
    outs = fn.apply(*xs)
    gin = torch.autograd.grad(outs, xs, [_t(grads[f.name]) for f in op.forward_output_fields])
    assert fn.forward_kernel.last_variant == 'march' and fn.backward_kernel.last_variant == 'march'
    ref_out, ref_din = forward_backward(op, ins, grads)
    for f, o in zip(op.forward_output_fields, outs):
        ref = ref_out[f.name]
        assert np.abs(o.detach().cpu().numpy() - ref).max() <= 1e-6 * max(np.abs(ref).max(), 1e-30), f.name
    for f, g_ in zip(op.forward_input_fields, gin):
        ref = ref_din['diff' + f.name]
        assert np.abs(g_.cpu().numpy() - ref).max() <= 1e-6 * max(np.abs(ref).max(), 1e-30), f.name


def test_timeloop_schedules_independent_calls_from_the_dependency_graph():
    """f-3: the step's parts are grouped into levels of mutually independent calls (ComputationGraph.levels over what each
    kernel reads / writes); the calls of a level go to different streams (parallel branches of the captured CUDA graph).
    Forward and adjoint of one step are independent, the two swaps follow; a third kernel reading the adjoint's output
    must wait for it.  Same bits as issuing everything in program order."""
    import torch
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    from pystencils_autodiff_b200.datahandling import SlabDataHandling
    shape = (64, 128)
    op = make_config('c2', shape=shape)
    fk, bk = CompiledKernel(op.forward_ast_gpu), CompiledKernel(op.backward_ast_gpu)
    w, diffu = ps.fields('w, diffu: float32[64,128]')
    third = CompiledKernel(ps.AutoDiffOp([ps.Assignment(w.center, 2 * diffu[0, 1] - diffu[1, 0])],
                                         boundary_handling='zeros').forward_ast_gpu)
    rng = np.random.default_rng(12)
    U0, G0 = rng.normal(size=shape).astype(np.float32), rng.normal(size=shape).astype(np.float32)
    finals = []
    for concurrent in (True, False):
        dh = SlabDataHandling(shape, 0, 1, 0, device='cuda:0')
        dh.add_arrays('u, out, diffout, diffu, w', dtype=np.float32)
        dh.owned('u').copy_(_t(U0))
        dh.owned('diffout').copy_(_t(G0))
        tl = dh.create_timeloop(use_cuda_graph=True, concurrent=concurrent)
        tl.add_call(fk, {})
        tl.add_call(bk, {})
        tl.add_call(third, {})
        tl.swap('u', 'out')
        assert tl.levels() == [[0, 1], [2, 3]]       # forward | adjoint, then the kernel reading diffu and the swap
        tl.run(9)
        torch.cuda.synchronize()
        finals.append({n: dh.owned(n).clone() for n in ('u', 'diffu', 'w')})
    for n in finals[0]:
        assert torch.equal(finals[0][n], finals[1][n]), n
    ref = evaluate(third.ir.assignments if hasattr(third.ir, 'assignments') else
                   [ps.Assignment(w.center, 2 * diffu[0, 1] - diffu[1, 0])], dict(diffu=finals[0]['diffu'].cpu().numpy()), 'zeros')['w']
    assert np.abs(finals[0]['w'].cpu().numpy() - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
    # an opaque step function makes the step unschedulable: program order
    tl2 = dh.create_timeloop()
    tl2.add_call(fk, {})
    tl2.add_single_step_function(lambda: None)
    assert tl2.levels() is None


@pytest.mark.no_launch          # the launches happen in the process this test starts
@pytest.mark.parametrize('name,steps', [('c3', 5), ('c4', 4)])
def test_periodic_time_loop_on_one_gpu(name, steps):
    """One rank on a periodic domain is its own neighbour (two device copies per synchronisation): single steps and fused
    pairs against a torch.roll restatement of the periodic stencil (``scripts/check_periodic.py`` as one process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'scripts', 'check_periodic.py'), name, str(steps)],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count('IDENTICAL') == 2 and 'DIFFERENT' not in out.stdout
