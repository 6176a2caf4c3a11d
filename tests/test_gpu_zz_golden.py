"""GPU: the torch_native Function (CUDA kernels through the C ABI) against the committed golden vectors — outputs and
gradients of the REFERENCE's own forward / backward assignments (executed from /root/reference in the build container by
tests/golden/make_reference_golden.py) on seeded float64 inputs.  Tolerance: north_star's 1e-12 relative for float64."""
import numpy as np
import pytest

from golden_util import build_op, golden_arrays, golden_names

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mode', [None, 'zeros'])
@pytest.mark.parametrize('name', golden_names())
def test_function_matches_reference_golden_vectors(name, mode):
    import torch
    op = build_op(name, mode)
    ins, outs, grads = golden_arrays(name, mode)
    fn = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    constant = {getattr(f, 'name', f) for f in op.constant_fields}        # the list also holds the name 'indexVector'
    tens = [torch.from_numpy(np.ascontiguousarray(ins[f.name])).cuda().requires_grad_(f.name not in constant)
            for f in op.forward_input_fields]
    res = fn.apply(*tens)
    assert isinstance(res, tuple) and len(res) == len(op.forward_output_fields)
    for f, t in zip(op.forward_output_fields, res):
        scale = max(1.0, np.abs(outs[f.name]).max())
        assert np.abs(t.detach().cpu().numpy() - outs[f.name]).max() <= 1e-12 * scale, (name, mode, f.name)
    upstream = [torch.from_numpy(np.ascontiguousarray(ins['diff' + f.name])).cuda() for f in op.forward_output_fields]
    torch.autograd.backward(res, upstream)
    checked = 0
    for f, t in zip(op.forward_input_fields, tens):
        key = 'diff' + f.name
        if key in grads and t.grad is not None:
            scale = max(1.0, np.abs(grads[key]).max())
            assert np.abs(t.grad.cpu().numpy() - grads[key]).max() <= 1e-12 * scale, (name, mode, key)
            checked += 1
    assert checked >= 1
    if name.endswith('_aligned'):
        assert fn.forward_kernel.last_variant == 'march' and fn.backward_kernel.last_variant == 'march'


def test_exact_adjoint_mode_passes_gradcheck_on_the_gpu():
    """``AutoDiffOp(..., adjoint_mode='exact')``: ``torch.autograd.gradcheck`` on random non-zero float64 inputs for a
    non-linear two-field stencil and for the TV gradient (C5), 'zeros' boundary.  The default mode restates the reference's
    rule (coefficients left at the centre cell), which is not the transpose there."""
    import sympy as sp
    import torch
    import pystencils_autodiff_b200 as ps
    from pystencils_autodiff_b200.configs import tv_gradient_op
    x, y, z = ps.fields('x, y, z: float64[7,8]')
    asg = ps.AssignmentCollection({z.center: x[1, 0] * y[0, 0] + sp.sin(x[0, -1]) * y[-1, 1]})
    gen = torch.Generator().manual_seed(0)
    for make_op, shape in ((lambda m: ps.AutoDiffOp(asg, op_name='nl_' + m, boundary_handling='zeros', adjoint_mode=m), (7, 8)),
                           (lambda m: tv_gradient_op(shape=(2, 6, 8), dtype='float64', adjoint_mode=m), (2, 6, 8))):
        ok = {}
        for mode in ('exact', 'reference'):
            fn = make_op(mode).create_tensorflow_op(backend='torch_native', use_cuda=True)
            ins = [(torch.rand(shape, generator=gen, dtype=torch.float64) + 0.5).cuda().requires_grad_(True)
                   for _ in fn.forward_ast.input_fields]
            ok[mode] = torch.autograd.gradcheck(fn.apply, ins, atol=1e-4, raise_exception=False)   # the reference tests' atol
        assert ok['exact'] is True
        assert ok['reference'] is False
