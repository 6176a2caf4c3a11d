"""GPU: the reference's time loop (kernel call, ghost-layer synchronisation, swap) on slabs, with pairs of steps fused
into one launch and one two-plane halo exchange per pair (SlabDataHandling.run_steps, datahandling.slab_ranges(steps=2)).
The CPU suite replays the same launches (tests/test_march_replay.py::test_slab_fused_steps_ranges_replay,
tests/test_slab_gloo.py::test_run_steps_fused_on_slabs)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import evaluate
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import heat3d_op, stencil27_op
from pystencils_autodiff_b200.datahandling import SlabDataHandling

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('make, tol', [(heat3d_op, 1e-6), (stencil27_op, 1e-12)])
@pytest.mark.parametrize('bh', ['zeros', None])
def test_run_steps_on_one_slab_with_ghost_planes(make, tol, bh):
    """One rank: the fused-pair kernel runs on a launch range (the owned planes of an array stored with two ghost planes
    per side) and equals both the oracle's time loop and the whole-array fused launches."""
    import torch
    shape, steps = (11, 30, 124), 5
    op = make(shape=shape, boundary_handling=bh)
    dt = op.forward_ast_gpu.input_fields[0].dtype.numpy_dtype
    u = np.random.default_rng(4).standard_normal(shape).astype(dt)
    ref = u
    for _ in range(steps):
        ref = evaluate(op.forward_assignments, {'u': ref}, boundary_handling=bh)['out'].astype(dt)
    dh = SlabDataHandling(shape, 0, 1, 2, 'cuda')
    dh.add_arrays('u, out', dtype=dt)
    op_l = make(shape=dh.dec.local_shape, boundary_handling=bh)
    kl = CompiledKernel(op_l.forward_ast_gpu)
    dh.owned('u').copy_(torch.from_numpy(u).cuda())
    res = dh.run_steps(kl, steps, fuse=True)
    assert [c for c in dh.call_queue if c[0] == 'KernelCall'] == [('KernelCall', kl.function_name, 2)] * 2 + \
        [('KernelCall', kl.function_name)]
    assert np.abs(res[dh.dec.owned].cpu().numpy() - ref).max() <= tol
    whole = CompiledKernel(op.forward_ast_gpu).run_steps(torch.from_numpy(u).cuda(), steps, fuse=True)
    assert torch.equal(res[dh.dec.owned], whole)
    ghosts = torch.cat([res[:2], res[-2:]])
    assert float(ghosts.abs().max()) == 0.0          # ghost planes are never written


@pytest.mark.no_launch          # the launches happen in the rank processes this test starts
@pytest.mark.parametrize('name,bh', [('c3', 'zeros'), ('c3', 'none'), ('c4', 'zeros')])
def test_fused_steps_sharded_equals_unsharded(name, bh):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2 if n < 4 else 4
    port = 29300 + (hash((name, bh)) % 200)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                          '--master-addr', '127.0.0.1', '--master-port', str(port),
                          os.path.join(ROOT, 'scripts', 'check_slab_steps.py'), name, bh, '5'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    # the ranks print concurrently: two reports can share a line, so count verdicts, not lines
    assert out.stdout.count('[rank') == 4 * world and out.stdout.count('IDENTICAL') == 4 * world, out.stdout[-2000:]
    assert 'DIFFERENT' not in out.stdout


@pytest.mark.parametrize('make, tol', [(heat3d_op, 1e-6), (stencil27_op, 1e-12)])
def test_slab_autograd_function_on_one_rank(make, tol):
    """``create_slab_autograd_function`` with a single rank and one ghost plane per side: a chain of two steps equals the
    plain torch_native Function bit for bit (outputs and gradients) and the oracle within tolerance; chained steps reuse
    the padded buffers."""
    import torch
    from pystencils_autodiff_b200.datahandling import create_slab_autograd_function
    shape = (12, 30, 124)
    op = make(shape=shape, boundary_handling='zeros')
    dt = op.forward_ast_gpu.input_fields[0].dtype.numpy_dtype
    rng = np.random.default_rng(9)
    U, R = rng.standard_normal(shape).astype(dt), rng.standard_normal(shape).astype(dt)
    dh = SlabDataHandling(shape, 0, 1, 1, 'cuda')
    Step = create_slab_autograd_function(op, dh)
    Plain = op.create_tensorflow_op(backend='torch_native', use_cuda=True)
    grads, outs = [], []
    for F in (Step, Plain):
        u = torch.from_numpy(U).cuda().requires_grad_(True)
        (o1,) = F.apply(u)
        (o2,) = F.apply(o1)
        (o2 * torch.from_numpy(R).cuda()).sum().backward()
        outs.append(o2.detach())
        grads.append(u.grad)
        if F is Step:
            assert o1._base is not None and o1._base.shape[0] == shape[0] + 2
    assert torch.equal(outs[0], outs[1]) and torch.equal(grads[0], grads[1])
    r1 = evaluate(op.forward_assignments, {'u': U.astype(np.float64)}, 'zeros')['out']
    r2 = evaluate(op.forward_assignments, {'u': r1}, 'zeros')['out']
    assert np.abs(outs[0].cpu().numpy() - r2).max() <= tol * max(1.0, np.abs(r2).max())
