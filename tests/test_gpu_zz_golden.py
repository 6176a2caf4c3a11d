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
    constant = {f.name for f in op.constant_fields}
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
