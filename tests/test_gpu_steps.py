"""GPU: pairs of unrolled steps fused into one launch (emit_chain.py, CompiledKernel.run_steps,
AutoDiffOp.create_unrolled_torch_op) against the oracle applied step by step."""
import numpy as np
import pytest

from oracle import evaluate
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import diffusion2d_op, heat3d_op, stencil27_op, tv_gradient_op

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _oracle_steps(assigns, fin, fout, u, steps, bh):
    r = u
    for _ in range(steps):
        r = evaluate(assigns, {fin: r}, boundary_handling=bh)[fout].astype(u.dtype)
    return r


@pytest.mark.parametrize('make, tol', [(heat3d_op, 1e-6), (stencil27_op, 1e-12)])
@pytest.mark.parametrize('bh', ['zeros', None])
@pytest.mark.parametrize('shape', [(19, 45, 252), (9, 30, 120), (5, 7, 8)])
def test_fused_pair_matches_oracle_twice(make, tol, bh, shape):
    """Ragged and tiny grids included: tiles overhang the array on every side, the intermediate field must be zero
    outside the iteration range of the first step.  Tolerances: north_star's 1e-6 (fp32) / 1e-12 (fp64)."""
    import torch
    op = make(shape=shape, boundary_handling=bh)
    for ir, assigns in ((op.forward_ast_gpu, op.forward_assignments), (op.backward_ast_gpu, op.backward_assignments)):
        k = CompiledKernel(ir)
        assert k.fused_steps_reason() is None
        fin, fout = ir.input_fields[0], ir.output_fields[0]
        u = np.random.default_rng(2).standard_normal(shape).astype(fin.dtype.numpy_dtype)
        ref = _oracle_steps(assigns, fin.name, fout.name, u, 2, bh)
        ut = _t(u)
        out = torch.full_like(ut, float('nan'))
        k(**{fin.name: ut, fout.name: out}, _variant='march_x2')
        assert k.last_instance == 'march_x2'
        assert np.abs(out.cpu().numpy() - ref).max() <= tol
        assert torch.equal(ut, _t(u))                     # the input is not touched


@pytest.mark.parametrize('steps', [1, 2, 3, 4, 7])
def test_run_steps_counts_and_parity(steps):
    """T steps = T // 2 fused launches + T % 2 single ones (fp32), ping-ponging through one scratch tensor."""
    from pystencils_autodiff_b200 import runtime
    shape = (12, 33, 124)
    op = heat3d_op(shape=shape)
    k = CompiledKernel(op.forward_ast_gpu)
    u = np.random.default_rng(steps).standard_normal(shape).astype(np.float32)
    ut = _t(u)
    n0 = runtime.launch_count()
    out = k.run_steps(ut, steps)
    assert runtime.launch_count() - n0 == steps // 2 + steps % 2
    ref = _oracle_steps(op.forward_assignments, 'u', 'out', u, steps, 'zeros')
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-6
    # fp64 fields: pairs too since round 2 (two CTAs per SM, exchanged rows: 1.11x at 768^3); single steps on request, and
    # both routes agree
    op64 = stencil27_op(shape=shape)
    k64 = CompiledKernel(op64.forward_ast_gpu)
    v = _t(np.random.default_rng(0).standard_normal(shape))
    n0 = runtime.launch_count()
    a = k64.run_steps(v, steps)
    assert runtime.launch_count() - n0 == steps // 2 + steps % 2
    n0 = runtime.launch_count()
    b = k64.run_steps(v, steps, fuse=False)
    assert runtime.launch_count() - n0 == steps
    assert (a - b).abs().max().item() <= 1e-13


def test_unrolled_function_forward_and_adjoint():
    """create_unrolled_torch_op(T): forward = S^T, backward = (S^T)^T — checked against T chained op.apply calls
    (what a user of the reference writes) and through the adjoint identity <S^T u, r> = <u, (S^T)' r>."""
    import torch
    shape = (10, 31, 124)
    op = heat3d_op(shape=shape)
    single = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    fused = op.create_unrolled_torch_op(5)
    rng = np.random.default_rng(7)
    u0 = rng.standard_normal(shape).astype(np.float32)
    r = _t(rng.standard_normal(shape).astype(np.float32))
    ua = _t(u0).requires_grad_(True)
    x = ua
    for _ in range(5):
        (x,) = single.apply(x)
    (x * r).sum().backward()
    ub = _t(u0).requires_grad_(True)
    (y,) = fused.apply(ub)
    (y * r).sum().backward()
    assert (x - y).abs().max().item() <= 1e-6
    assert (ua.grad - ub.grad).abs().max().item() <= 1e-6
    lhs = (y.double() * r.double()).sum().item()
    rhs = (ub.detach().double() * ub.grad.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))
    # fp64: gradcheck-grade agreement with finite differences of the T-step map (linear, so FD is exact up to rounding)
    op64 = stencil27_op(shape=(6, 9, 16))
    f64 = op64.create_unrolled_torch_op(4)
    w = _t(rng.standard_normal((6, 9, 16))).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda t: f64.apply(t)[0], (w,), eps=1e-6, atol=1e-9)


def test_fused_steps_error_behaviour():
    import torch
    op = tv_gradient_op(shape=(2, 16, 32))
    with pytest.raises(ValueError, match='one input and one output'):
        op.create_unrolled_torch_op(2)
    k = CompiledKernel(heat3d_op(shape=(8, 16, 32)).forward_ast_gpu)
    u = torch.zeros((8, 16, 32), device='cuda')
    with pytest.raises(ValueError, match='different tensors'):
        k(u=u, out=u, _variant='march_x2')
    # a launch range is accepted for 3-D fused pairs (slab launches): whole-array range == no range, bit for bit
    rng = np.random.default_rng(3)
    u.copy_(torch.from_numpy(rng.standard_normal((8, 16, 32)).astype(np.float32)))
    whole, ranged = torch.empty_like(u), torch.empty_like(u)
    k(u=u, out=whole, _variant='march_x2')
    k(u=u, out=ranged, _variant='march_x2',
      _range=dict(iter_lo=[0, 0, 0], iter_hi=[8, 16, 32], write_lo=[0, 0, 0], write_hi=[8, 16, 32]))
    assert torch.equal(whole, ranged)
    # ... and still rejected for 2-D kernels lifted to one plane
    k2 = CompiledKernel(diffusion2d_op(shape=(16, 32), boundary_handling='zeros').forward_ast_gpu)
    u2 = torch.zeros((16, 32), device='cuda')
    with pytest.raises(ValueError, match='whole arrays'):
        k2(u=u2, out=torch.empty_like(u2), _variant='march_x2',
           _range=dict(iter_lo=[0, 0], iter_hi=[16, 32], write_lo=[0, 0], write_hi=[16, 32]))
    ut = torch.zeros((8, 16, 34), device='cuda')[:, :, 1:33]           # rows not 16-byte aligned -> generic kernel only
    with pytest.raises(ValueError, match='cannot be fused'):
        CompiledKernel(heat3d_op(shape=(8, 16, 32)).forward_ast_gpu).run_steps(ut, 2, fuse=True)
    out = CompiledKernel(heat3d_op(shape=(8, 16, 32)).forward_ast_gpu).run_steps(ut, 2)
    assert out.shape == ut.shape
