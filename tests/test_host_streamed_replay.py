"""``HostStreamedOp`` (bench.py's ``e2e`` leg at N = 1: host-resident fields streamed through the GPU in chunks of planes on
three streams) executed on CPU tensors: stand-in streams / events (tests/fake_cuda.py), the emitted kernels replayed as
the launches (tests/replay_kernels.py).  Checks the chunk / halo / launch-range bookkeeping against the oracle on the
whole field; the GPU suite checks the real thing (tests/test_gpu_parity.py::test_host_streamed_equals_resident)."""
import numpy as np
import pytest
import torch

from fake_cuda import fake_cuda
from oracle import forward_backward
from pystencils_autodiff_b200 import configs
from pystencils_autodiff_b200.datahandling import HostStreamedOp
from replay_kernels import ReplayKernel


@pytest.mark.parametrize('make, shape, bh, chunk, tol', [
    (configs.heat3d_op, (23, 10, 132), 'zeros', 6, 2e-6),         # 4 chunks, the last one partial, ring of 3 buffers reused
    (configs.heat3d_op, (17, 10, 132), None, 5, 2e-6),            # interior iteration: global border clipped per chunk
    (configs.stencil27_op, (11, 9, 36), 'zeros', 4, 1e-13),
    (configs.tv_gradient_op, (5, 12, 36), 'zeros', 2, 2e-5),      # no reach along dim 0: chunks without ghost planes
])
@pytest.mark.parametrize('ramp', [True, False])
def test_host_streamed_chunks_equal_whole_field(make, shape, bh, chunk, tol, ramp):
    op = make(shape=shape, boundary_handling=bh)
    rng = np.random.default_rng(3)
    ins = {f.name: rng.uniform(0.1, 1.0, shape).astype(f.dtype.numpy_dtype) for f in op.forward_input_fields}
    grads = {f.name: rng.standard_normal(shape).astype(f.dtype.numpy_dtype) for f in op.forward_output_fields}
    with fake_cuda():
        streamed = HostStreamedOp(op, shape, device='cpu', chunk_planes=chunk, stages=3, ramp=ramp)
        streamed.fwd, streamed.bwd = ReplayKernel(op.forward_ast_gpu), ReplayKernel(op.backward_ast_gpu)
        assert sum(streamed.sizes) == shape[0] and max(streamed.sizes) <= chunk and min(streamed.sizes) >= 1
        assert streamed.starts == [sum(streamed.sizes[:i]) for i in range(streamed.n_chunks)]
        if not ramp:
            assert streamed.n_chunks == -(-shape[0] // chunk)
        host_in = {n: torch.from_numpy(ins[n]) for n in ins}
        host_in.update({'diff' + n: torch.from_numpy(g) for n, g in grads.items()})
        host_out = {n: torch.full(shape, float('nan'), dtype=host_in[next(iter(ins))].dtype) for n in streamed.output_names}
        assert sorted(host_in) == sorted(streamed.input_names)
        streamed(host_in, host_out)
    ref_out, ref_grads = forward_backward(op, ins, grads)
    for name, ref in list(ref_out.items()) + list(ref_grads.items()):
        got = host_out[name].numpy()
        assert np.isfinite(got).all(), name
        assert np.abs(got - ref).max() <= tol * max(1.0, np.abs(ref).max()), name
    # every input plane crosses PCIe exactly once: the planes two neighbouring chunks share are copied on the device
    per_plane = int(np.prod(shape[1:])) * host_in[next(iter(ins))].element_size()
    assert streamed.h2d_bytes == shape[0] * per_plane * len(streamed.input_names)
    assert streamed.d2h_bytes == shape[0] * per_plane * len(streamed.output_names)


def test_chunk_sizes_ramp():
    """Short chunks at both ends (pipeline fill / drain), full ones in between; short fields fall back to uniform cuts."""
    f = HostStreamedOp._chunk_sizes
    c3 = f(1024, 48, 2)                                # the C3 field: 48-plane chunks of 4 MiB planes
    assert sum(c3) == 1024 and c3[:2] == [12, 24] and c3[-2:] == [24, 12] and set(c3[2:-2]) <= {47, 48}
    assert f(23, 6, 2) == [2, 3, 5, 4, 4, 3, 2]
    assert f(10, 6, 2) == [6, 4]                       # too short for a ramp
    assert f(16, 64, 2) == [16]
    for n0, c, m in ((768, 40, 2), (100, 7, 4), (5, 2, 1), (4096, 48, 1)):
        sizes = f(n0, c, m)
        assert sum(sizes) == n0 and max(sizes) <= c and all(v >= 1 for v in sizes)


def test_host_streamed_scalar_parameter_and_accumulate_form():
    """The CPU twin of tests/test_gpu_api.py::test_host_streamed_op_with_scalar_and_accumulate_form."""
    import sympy as sp
    import pystencils_autodiff_b200 as ps
    shape = (21, 16, 64)
    u, out = ps.fields('u, out: float64[%d,%d,%d]' % shape)
    al = sp.Symbol('alpha')
    fa = [ps.Assignment(out.center, u[0, 0, 0] + al * (u[1, 0, 0] + u[-1, 0, 0] + u[0, 0, 1] - 3 * u[0, 0, 0]))]
    op = ps.AutoDiffOp(fa, boundary_handling='zeros', time_constant_fields=[u])
    rng = np.random.default_rng(21)
    U, G = rng.normal(size=shape), rng.normal(size=shape)
    host = {'u': torch.from_numpy(U), 'diffout': torch.from_numpy(G),
            'out': torch.full(shape, float('nan'), dtype=torch.float64), 'diffu': torch.full(shape, float('nan'), dtype=torch.float64)}
    with fake_cuda():
        st = HostStreamedOp(op, shape, 'cpu', chunk_planes=4)
        st.fwd, st.bwd = ReplayKernel(op.forward_ast_gpu), ReplayKernel(op.backward_ast_gpu)
        with pytest.raises(TypeError):
            st({n: host[n] for n in st.input_names}, {n: host[n] for n in st.output_names})
        st({n: host[n] for n in st.input_names}, {n: host[n] for n in st.output_names}, alpha=0.3)
    ref_o, ref_d = forward_backward(op, dict(u=U), dict(out=G), scalars=dict(alpha=0.3))
    np.testing.assert_allclose(host['out'].numpy(), ref_o['out'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(host['diffu'].numpy(), ref_d['diffu'], rtol=1e-12, atol=1e-12)


def test_end_to_end_leg_one_rank_checks_itself():
    """``SlabStencilOp.end_to_end`` at N = 1 (bench.py's ``e2e``): after the timed steps a few planes of every streamed
    output are compared bit for bit with the resident kernels; the line carries ``matches_resident``."""
    from pystencils_autodiff_b200.datahandling import SlabStencilOp
    shape = (40, 10, 132)
    op = configs.heat3d_op(shape=shape, boundary_handling='zeros')
    with fake_cuda():
        slab = SlabStencilOp(op, shape, 0, 1, device='cpu', backend='torch')
        slab.fwd, slab.bwd = ReplayKernel(op.forward_ast_gpu), ReplayKernel(op.backward_ast_gpu)
        slab.randomize(torch.Generator().manual_seed(0))
        slab._fn = HostStreamedOp(op, shape, 'cpu', chunk_planes=8)
        slab._fn.fwd, slab._fn.bwd = ReplayKernel(op.forward_ast_gpu), ReplayKernel(op.backward_ast_gpu)
        slab._host = {n: torch.empty(shape, dtype=slab.dh.gpu_arrays[n].dtype) for n in slab._fn.fields}
        for n in slab._fn.input_names:
            slab._host[n].copy_(slab.dh.owned(n))
        r = slab.end_to_end(1, lambda: None)
        assert r['matches_resident'] is True
        assert r['checked_planes'][0] == 0 and r['checked_planes'][-1] == shape[0] - 1 and len(r['checked_planes']) >= 4
        assert r['h2d'] == 2 * 4 * int(np.prod(shape)) and r['d2h'] == r['h2d']
        # a corrupted download is noticed
        slab._fn.output_names = list(slab._fn.output_names)
        real_call = HostStreamedOp.__call__

        def corrupting(self, host_in, host_out, **kw):
            real_call(self, host_in, host_out, **kw)
            host_out['out'][0, 0, 0] += 1.0
        HostStreamedOp.__call__ = corrupting
        try:
            assert slab.end_to_end(1, lambda: None)['matches_resident'] is False
        finally:
            HostStreamedOp.__call__ = real_call
