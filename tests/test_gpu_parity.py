"""GPU parity: CUDA kernels (through the torch_native Function and the C ABI) vs the numpy oracle.

Tolerances (BASELINE.json north_star): norm-wise relative error <= 1e-6 for float32 fields, <= 1e-12 for float64.
"""
import numpy as np
import pytest

from oracle import forward_backward
from pystencils_autodiff_b200.configs import make_config

pytestmark = pytest.mark.gpu

TOL = {'float32': 1e-6, 'float64': 1e-12}


def _rel(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


def run_op(op, shape, lo, hi, seed, variant=None):
    import torch
    dev = torch.device('cuda:0')
    rng = np.random.default_rng(seed)
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    ins = {f.name: rng.uniform(lo, hi, size=shape).astype(f.dtype.numpy_dtype) for f in op.forward_input_fields}
    grads = {f.name: rng.normal(size=shape).astype(f.dtype.numpy_dtype) for f in op.forward_output_fields}
    tens = [torch.from_numpy(ins[f.name]).to(dev).requires_grad_(True) for f in op.forward_input_fields]
    if variant is not None:
        fn.forward_kernel._select_variant = lambda tensors: variant
        fn.backward_kernel._select_variant = lambda tensors: variant
    outs = fn.apply(*tens)
    torch.autograd.backward(outs, [torch.from_numpy(grads[f.name]).to(dev) for f in op.forward_output_fields])
    ref_out, ref_din = forward_backward(op, ins, grads)
    res = {}
    for f, o in zip(op.forward_output_fields, outs):
        res[f.name] = _rel(o.detach().cpu().numpy(), ref_out[f.name])
    for f, t in zip(op.forward_input_fields, tens):
        res['diff' + f.name] = _rel(t.grad.cpu().numpy(), ref_din['diff' + f.name])
    return res, fn


CASES = [
    # name, shape, dtype, value range
    ('c1', (20, 30), 'float32', (0.5, 1.5)),
    ('c1', (20, 32), 'float64', (0.5, 1.5)),
    ('c2', (96, 256), 'float32', (-1, 1)),
    ('c2', (70, 132), 'float32', (-1, 1)),     # ragged: partial tiles in both directions
    ('c2', (33, 20), 'float64', (-1, 1)),
    ('c3', (40, 48, 256), 'float32', (-1, 1)),
    ('c3', (19, 21, 36), 'float32', (-1, 1)),   # ragged
    ('c3', (9, 10, 11), 'float32', (-1, 1)),    # unaligned pitch -> generic
    ('c4', (12, 20, 132), 'float64', (-1, 1)),
    ('c5', (3, 40, 136), 'float32', (0, 1)),
]


@pytest.mark.parametrize('bh', [None, 'zeros'])
@pytest.mark.parametrize('name,shape,dtype,rng', CASES)
def test_forward_backward_matches_oracle(name, shape, dtype, rng, bh):
    op = make_config(name, shape=shape, dtype=dtype, boundary_handling=bh)
    res, fn = run_op(op, shape, rng[0], rng[1], seed=1)
    tol = TOL[dtype]      # north_star: 1e-6 (fp32) / 1e-12 (fp64), the TV gradient's sqrt / division chains included
    for k, v in res.items():
        assert v <= tol, (name, shape, bh, k, v, fn.forward_kernel.last_variant, fn.backward_kernel.last_variant)


@pytest.mark.parametrize('bh', [None, 'zeros'])
@pytest.mark.parametrize('name,shape,dtype,rng', [c for c in CASES if c[0] != 'c1'])
def test_generic_variant_matches_oracle(name, shape, dtype, rng, bh):
    op = make_config(name, shape=shape, dtype=dtype, boundary_handling=bh)
    res, fn = run_op(op, shape, rng[0], rng[1], seed=2, variant='generic')
    tol = TOL[dtype]
    for k, v in res.items():
        assert v <= tol, (name, shape, bh, k, v)


def test_march_variant_is_used_for_aligned_shapes():
    op = make_config('c3', shape=(16, 32, 128), dtype='float32')
    res, fn = run_op(op, (16, 32, 128), -1, 1, seed=3)
    assert fn.forward_kernel.last_variant == 'march'
    assert fn.backward_kernel.last_variant == 'march'


@pytest.mark.parametrize('name,shape', [('c2', (12, 16)), ('c3', (6, 6, 8)), ('c4', (5, 6, 8))])
def test_gradcheck_zeros_boundary(name, shape):
    """Same gate as the reference (tests/test_tfmad.py:186-231: gradcheck, atol=1e-4, float64, 'zeros') but on
    random non-zero inputs."""
    import torch
    op = make_config(name, shape=shape, dtype='float64', boundary_handling='zeros')
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    torch.manual_seed(0)
    tens = tuple(torch.randn(*shape, dtype=torch.float64, device='cuda', requires_grad=True)
                 for _ in op.forward_input_fields)
    assert torch.autograd.gradcheck(fn.apply, tens, atol=1e-4, raise_exception=True)


@pytest.mark.parametrize('name,shape,chunk', [('c3', (23, 24, 128), 5), ('c4', (11, 16, 64), 4), ('c5', (5, 16, 128), 2),
                                              ('c2', (70, 128), 16)])
@pytest.mark.parametrize('bh', [None, 'zeros'])
def test_host_streamed_equals_resident(name, shape, chunk, bh):
    """Chunked H2D / compute / D2H pipeline == whole-field kernels, bit for bit."""
    import torch
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    from pystencils_autodiff_b200.datahandling import HostStreamedOp
    op = make_config(name, shape=shape, boundary_handling=bh)
    st = HostStreamedOp(op, shape, 'cuda:0', chunk_planes=chunk)
    assert st.n_chunks > 2
    torch.manual_seed(0)
    host = {n: (torch.rand(shape, dtype=getattr(torch, f.dtype.numpy_dtype.name)) + 0.1).pin_memory()
            for n, f in st.fields.items()}
    st({n: host[n] for n in st.input_names}, {n: host[n] for n in st.output_names})
    torch.cuda.synchronize()
    dev = {n: host[n].cuda() for n in st.input_names}
    for kern in (CompiledKernel(op.forward_ast_gpu), CompiledKernel(op.backward_ast_gpu)):
        for f in kern.ir.output_fields:
            dev[f.name] = torch.empty(shape, dtype=host[f.name].dtype, device='cuda:0')
        kern(**{f.name: dev[f.name] for f in kern.fields})
    torch.cuda.synchronize()
    for n in st.output_names:
        assert torch.equal(host[n], dev[n].cpu()), n


def _run_raw(asg, arrays_np, bh=None, scalars=None):
    """Kernel call path (``CompiledKernel(**tensors)``) vs oracle for arbitrary assignment collections."""
    import torch
    import pystencils_autodiff_b200 as ps
    from oracle import evaluate
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    from pystencils_autodiff_b200.ir import lower_assignments
    ir = lower_assignments(asg, bh, 'k')
    k = CompiledKernel(ir)
    ref = evaluate(asg, arrays_np, bh, scalars)
    tens = {n: torch.from_numpy(np.ascontiguousarray(a)).cuda() for n, a in arrays_np.items()}
    for f in ir.output_fields:
        if f.name not in tens:
            tens[f.name] = torch.full(ref[f.name].shape, float('nan'), dtype=getattr(torch, ref[f.name].dtype.name),
                                      device='cuda')
    k(**{f.name: tens[f.name] for f in k.fields}, **(scalars or {}))
    return {n: tens[n].cpu().numpy() for n in ref}, ref, k


@pytest.mark.parametrize('bh', [None, 'zeros'])
def test_one_dimensional_field(bh):
    import pystencils_autodiff_b200 as ps
    a, b = ps.fields('a, b: float64[257]')
    asg = ps.AssignmentCollection({b.center: 0.25 * a[-2] - a[1] + a[0] ** 2})
    got, ref, k = _run_raw(asg, dict(a=np.random.default_rng(0).normal(size=257)), bh)
    assert k.last_variant == 'generic'
    np.testing.assert_allclose(got['b'], ref['b'], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize('bh', [None, 'zeros'])
def test_index_dimension_fields(bh):
    """Vector fields (one index dimension): the curl of tests/test_tfmad.py:341-401 and its adjoint."""
    import pystencils_autodiff_b200 as ps
    u = ps.Field.create_fixed_size('curl_input', (20, 30), index_dimensions=0)
    c = ps.Field.create_fixed_size('curl', (20, 30, 2), index_dimensions=1)
    disc = ps.fd.Discretization2ndOrder(dx=1)
    fa = ps.AssignmentCollection([ps.Assignment(c.center(0), disc(ps.fd.Diff(u, 0))),
                                  ps.Assignment(c.center(1), disc(ps.fd.Diff(u, 1)))], [])
    op = ps.AutoDiffOp(fa, boundary_handling=bh)
    rng = np.random.default_rng(1)
    got, ref, k = _run_raw(op.forward_assignments, dict(curl_input=rng.normal(size=(20, 30))), bh)
    assert k.last_variant == 'generic'
    np.testing.assert_allclose(got['curl'], ref['curl'], rtol=1e-13, atol=1e-13)
    got, ref, k = _run_raw(op.backward_assignments, dict(diffcurl=rng.normal(size=(20, 30, 2))), bh)
    np.testing.assert_allclose(got['diffcurl_input'], ref['diffcurl_input'], rtol=1e-13, atol=1e-13)


def test_off_centre_write_and_wide_offsets():
    import pystencils_autodiff_b200 as ps
    x, y = ps.fields('x, y: float32[40,136]')
    asg = ps.AssignmentCollection({y[0, 1]: x[0, 0] + 2 * x[-1, 0]})            # off-centre lhs -> generic
    got, ref, k = _run_raw(asg, dict(x=np.random.default_rng(2).normal(size=(40, 136)).astype(np.float32)))
    assert k.last_variant == 'generic'
    np.testing.assert_allclose(got['y'], ref['y'], rtol=1e-6, atol=1e-6)
    asg = ps.AssignmentCollection({y.center: x[0, 3] - x[0, -4] + x[2, 0] * x[-3, 1]})   # |dx| up to the strip width
    for bh in (None, 'zeros'):
        got, ref, k = _run_raw(asg, dict(x=np.random.default_rng(3).normal(size=(40, 136)).astype(np.float32)), bh)
        assert k.last_variant == 'march'
        np.testing.assert_allclose(got['y'], ref['y'], rtol=2e-6, atol=2e-6)
    asg = ps.AssignmentCollection({y.center: x[0, 5] + x[0, 0]})                # wider than the strip -> generic
    got, ref, k = _run_raw(asg, dict(x=np.random.default_rng(4).normal(size=(40, 136)).astype(np.float32)), 'zeros')
    assert k.last_variant == 'generic'
    np.testing.assert_allclose(got['y'], ref['y'], rtol=1e-6, atol=1e-6)


def test_asymmetric_3d_stencil_with_two_inputs_and_scalar():
    """Asymmetric z halo (0 below, 2 above), two input fields with different halos, a scalar parameter, fp64."""
    import sympy as sp
    import pystencils_autodiff_b200 as ps
    a, b, o = ps.fields('a, b, o: float64[9,12,64]')
    s = sp.Symbol('s')
    asg = ps.AssignmentCollection({o.center: s * a[2, 0, 0] - a[1, -1, 1] * b[0, 0, 0] + sp.exp(-b[0, 1, 0] ** 2) + a[0, 0, -2]})
    rng = np.random.default_rng(5)
    arrays = dict(a=rng.normal(size=(9, 12, 64)), b=rng.normal(size=(9, 12, 64)))
    for bh in (None, 'zeros'):
        got, ref, k = _run_raw(asg, arrays, bh, scalars=dict(s=0.75))
        assert k.last_variant == 'march'
        np.testing.assert_allclose(got['o'], ref['o'], rtol=1e-12, atol=1e-12)
        op = ps.AutoDiffOp(asg, boundary_handling=bh)
        env = dict(arrays, diffo=rng.normal(size=(9, 12, 64)))
        got, ref, k = _run_raw(op.backward_assignments, env, bh, scalars=dict(s=0.75))
        for n in ref:
            np.testing.assert_allclose(got[n], ref[n], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('bh', [None, 'zeros'])
def test_index_dimension_fields_soa_layout_use_the_fast_path(bh):
    """Vector fields stored structure-of-arrays (x contiguous, like the reference's 'fzyx' GPU layout) are split into
    per-component scalar fields and run through the TMA march kernels."""
    import torch
    import pystencils_autodiff_b200 as ps
    from oracle import evaluate
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    shape = (24, 128)
    u = ps.Field.create_fixed_size('curl_input', shape, index_dimensions=0, dtype=np.float32)
    c = ps.Field.create_fixed_size('curl', shape + (2,), index_dimensions=1, dtype=np.float32)
    disc = ps.fd.Discretization2ndOrder(dx=1)
    fa = ps.AssignmentCollection([ps.Assignment(c.center(0), disc(ps.fd.Diff(u, 0)) + 0.5 * u.center),
                                  ps.Assignment(c.center(1), disc(ps.fd.Diff(u, 1)))], [])
    op = ps.AutoDiffOp(fa, boundary_handling=bh)
    rng = np.random.default_rng(8)
    U = rng.normal(size=shape).astype(np.float32)
    DC = rng.normal(size=shape + (2,)).astype(np.float32)

    def soa(a):          # [y, x, f] view of an [f, y, x]-contiguous buffer
        t = torch.from_numpy(np.ascontiguousarray(np.moveaxis(a, -1, 0))).cuda()
        return t.permute(1, 2, 0)

    fk, bk = CompiledKernel(op.forward_ast_gpu), CompiledKernel(op.backward_ast_gpu)
    curl = soa(np.full(shape + (2,), np.nan, dtype=np.float32))
    fk(curl_input=torch.from_numpy(U).cuda(), curl=curl)
    assert fk.last_variant == 'march'
    ref = evaluate(op.forward_assignments, dict(curl_input=U), bh)['curl']
    assert np.abs(curl.cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()
    du = torch.full(shape, float('nan'), device='cuda')
    bk(diffcurl=soa(DC), diffcurl_input=du)
    assert bk.last_variant == 'march'
    ref = evaluate(op.backward_assignments, dict(diffcurl=DC), bh)['diffcurl_input']
    assert np.abs(du.cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()
    # the same tensors in array-of-structs layout fall back to the generic kernel and agree
    curl2 = torch.full(shape + (2,), float('nan'), device='cuda')
    fk(curl_input=torch.from_numpy(U).cuda(), curl=curl2)
    assert fk.last_variant == 'generic'
    assert (curl2 - curl).abs().max().item() <= 1e-6
