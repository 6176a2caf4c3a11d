"""Emitter: determinism, window analysis and the structure of the specialised kernels."""
import pytest

from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import make_config
from pystencils_autodiff_b200.emit import MarchTuning, emit_kernel, emit_march, march_ineligible_reason
import pystencils_autodiff_b200 as ps


def test_codegen_is_reproducible():
    """tests/backends/test_torch_native_compilation.py:214-244: identical source over 10 runs, sympy cache cleared."""
    from sympy.core.cache import clear_cache
    first = None
    for _ in range(10):
        op = make_config('c5', shape=(2, 16, 128))
        fn_src = CompiledKernel(op.forward_ast_gpu).code + CompiledKernel(op.backward_ast_gpu).code
        keys = (CompiledKernel(op.forward_ast_gpu).emitted('march').cache_key,
                CompiledKernel(op.backward_ast_gpu).emitted('march').cache_key)
        clear_cache()
        if first is None:
            first = (fn_src, keys)
        assert (fn_src, keys) == first


@pytest.mark.parametrize('name', ['c1', 'c2', 'c3', 'c4', 'c5'])
def test_every_config_emits_both_variants(name):
    op = make_config(name)
    for ir in (op.forward_ast_gpu, op.backward_ast_gpu):
        k = CompiledKernel(ir)
        assert k.variants == ['generic', 'march']
        src = k.emitted('march').source
        assert 'psad_march.cuh' in src and 'psad_lds_vec' in src and 'psad_stg_vec' in src
        assert k.emitted('march').options == ['-fmad=false']


def test_bytes_per_cell_match_the_roofline_table():
    """BASELINE.md section 2."""
    expect = {'c1': (12, 20), 'c2': (8, 8), 'c3': (8, 8), 'c4': (16, 16), 'c5': (12, 16)}
    for name, (f, b) in expect.items():
        op = make_config(name)
        assert op.forward_ast_gpu.bytes_per_cell() == f
        assert op.backward_ast_gpu.bytes_per_cell() == b


def test_seven_point_window():
    """Each staged element is read from shared memory once: the centre rows of plane z+1 arrive as 128-bit loads,
    move through the window (no further loads at z and z-1), rows y+-1 are loaded at plane z only."""
    op = make_config('c3', shape=(16, 32, 128))
    ek = emit_march(op.forward_ast_gpu, MarchTuning(ty=32, ry=2))
    body = ek.source.split('psad_step_ph0')[1].split('psad_step_ph1')[0]
    assert body.count('psad_lds_vec<float>') == 4          # 2 centre rows (z+1) + rows y-1, y+2 (z)
    assert body.count('psad_from_left') == 2 and body.count('psad_from_right') == 2
    assert 'cfg::NP' in open(__import__('os').path.join(
        __import__('os').path.dirname(ps.__file__), 'csrc', 'kernels', 'psad_march.cuh')).read()
    assert 'NP = 3' in ek.source and 'JREL = 1' in ek.source
    nocarry = emit_march(op.forward_ast_gpu, MarchTuning(ty=32, ry=2, carry=False))
    assert 'NP = 1' in nocarry.source and 'JREL = 0' in nocarry.source


def test_sum_order_is_independent_of_window_phase():
    """The three phase instances must round identically: same FMA chain order, only register names rotate."""
    import re
    op = make_config('c4', shape=(8, 16, 64))
    ek = emit_march(op.forward_ast_gpu)
    bodies = [ek.source.split('psad_step_ph%d(' % p)[1].split('PSAD_DEV void')[0] for p in range(3)]
    exprs = [[l for l in b.splitlines() if 'o0[0] =' in l][0] for b in bodies]
    norm = {re.sub(r'_k\d', '_k', e) for e in exprs}
    assert len(norm) == 1
    assert len(set(exprs)) == 3


def test_march_eligibility():
    u, out = ps.fields('u, out: float32[8,128]')
    assert march_ineligible_reason(ps.AutoDiffOp([ps.Assignment(out.center, u[0, 1])]).forward_ast_gpu) is None
    ir = ps.AutoDiffOp([ps.Assignment(out[0, 1], u[0, 0])]).forward_ast_gpu
    assert march_ineligible_reason(ir) == 'off-centre writes'
    assert emit_kernel(ir).kind == 'generic'
    v, w = ps.fields('v(2), w: float32[8,128]')
    ir = ps.AutoDiffOp([ps.Assignment(w.center, v.center(0) + v[0, 1](1))]).forward_ast_gpu
    assert march_ineligible_reason(ir) == 'index dimensions'
    a, b = ps.fields('a, b: float64[128]')
    assert 'spatial' in march_ineligible_reason(ps.AutoDiffOp([ps.Assignment(b.center, a[1])]).forward_ast_gpu)


def test_iteration_space_rule():
    u, out = ps.fields('u, out: float32[16,128]')
    asg = [ps.Assignment(out.center, u[2, 0] + u[0, -1])]
    assert ps.AutoDiffOp(asg).forward_ast_gpu.ghost_layers == 2
    assert ps.AutoDiffOp(asg, boundary_handling='zeros').forward_ast_gpu.ghost_layers == 0
    assert ps.AutoDiffOp(asg, boundary_handling='zeros').forward_ast_gpu.boundary == 'zeros'
    # the symbolic ConditionalFieldAccess form lowers to the same IR
    from pystencils_autodiff_b200.ir import lower_assignments
    ir = lower_assignments(ps.add_fixed_constant_boundary_handling(ps.AssignmentCollection(asg)), None)
    assert ir.boundary == 'zeros' and ir.ghost_layers == 0
    assert ir.halo('u') == [(0, 2), (1, 0)]


def test_show_code_and_kernel_options():
    op = make_config('c5', shape=(2, 16, 128), fast_math=True)
    k = CompiledKernel(op.forward_ast_gpu)
    assert k.emitted('march').options == ['-fmad=false', '-ftz=true', '-prec-div=false', '-prec-sqrt=false']
    assert 'psad_rsqrt(' in k.emitted('march').source and 'powf' not in k.emitted('march').source
    src = ps.show_code(op.forward_ast_gpu)
    assert 'psad_tvgrad_forward_gpu' in src and 'PSAD_KERNEL_NAME' in src
    # a mask-free instance exists for every march kernel and differs only in the selects
    masked, plain = k.emitted('march').source, k.emitted('march_nomask').source
    assert '? (float)(' in masked and '? (float)(' not in plain


def test_tiles_fall_back_when_the_ring_does_not_fit():
    import sympy as sp
    a, b, o = ps.fields('a, b, o: float64[9,12,64]')
    asg = ps.AssignmentCollection({o.center: a[2, 0, 0] - a[1, -1, 1] * b[0, 0, 0] + sp.exp(-b[0, 1, 0] ** 2) + a[0, 0, -2]})
    op = ps.AutoDiffOp(asg, boundary_handling='zeros')
    for ir in (op.forward_ast_gpu, op.backward_ast_gpu):
        ek = emit_march(ir)
        assert ek.plan['smem_bytes'] <= 227 * 1024
        assert ek.plan['threads'] % 128 == 0        # 4k-1 consumer warps + the producer warp


def test_linear_plan_shares_partial_sums():
    from pystencils_autodiff_b200.linopt import plan_linear
    import sympy as sp
    U = {(r, c): sp.Symbol('u_%d_%d' % (r + 1, c + 1)) for r in range(-1, 3) for c in range(-1, 5)}

    def plane(r, c, wc, wf, wk):
        return (wc * U[r, c] + wf * (U[r - 1, c] + U[r + 1, c] + U[r, c - 1] + U[r, c + 1])
                + wk * (U[r - 1, c - 1] + U[r - 1, c + 1] + U[r + 1, c - 1] + U[r + 1, c + 1]))
    targets = []
    for r in range(2):
        for c in range(4):
            targets += [(('q', r, c), plane(r, c, 0.05, 0.02, 0.0075)), (('p', r, c), plane(r, c, 0.4, 0.05, 0.02))]
    plan = plan_linear(targets, set(U.values()))
    assert plan.op_count() <= 15 * 8 < 34 * 8          # 34 flops per cell evaluated naively
    # the plan is an exact re-association: evaluate it numerically
    import random
    random.seed(0)
    vals = {str(s): random.random() for s in U.values()}
    env = dict(vals)
    for nm, x, y in plan.temps:
        env[nm] = env[x] + env[y]
    for nm, adds in plan.sums:
        env[nm] = sum(env[x] for x in adds)
    for (key, expr), (key2, lst) in zip(targets, plan.targets):
        assert key == key2
        got = sum(float(cw) * env[nm] for cw, nm in lst)
        ref = float(expr.subs({s: vals[str(s)] for s in U.values()}))
        assert abs(got - ref) < 1e-12
    assert plan_linear([(0, U[0, 0] * U[0, 1])], set(U.values())) is None      # not linear -> no plan


def test_create_forward_and_backward_kernel_like_the_reference():
    """``AutoDiffOp.create_forward_kernel`` / ``create_backward_kernel`` (reference _autodiff.py:592-598:
    ``ps.create_kernel(assignments, *args, **kwargs).compile()`` on the raw assignments): GPU kernels with the given
    ``ghost_layers`` — no boundary transform even when the op has one — and pystencils' default target 'cpu' raises."""
    import pytest
    from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
    from pystencils_autodiff_b200.configs import make_config
    op = make_config('c2', shape=(16, 32), boundary_handling='zeros')
    fk = op.create_forward_kernel(target='gpu')
    assert isinstance(fk, CompiledKernel) and fk.ir.boundary == 'none' and fk.ir.ghost_layers == 1
    assert [f.name for f in fk.ir.input_fields] == ['u'] and [f.name for f in fk.ir.output_fields] == ['out']
    bk = op.create_backward_kernel('gpu', ghost_layers=2)
    assert bk.ir.ghost_layers == 2 and [f.name for f in bk.ir.input_fields] == ['diffout']
    assert op.forward_ast_gpu.boundary == 'zeros'            # the op's own kernels are untouched
    with pytest.raises(NotImplementedError, match='no CPU'):
        op.create_forward_kernel()
    with pytest.raises(NotImplementedError, match='no CPU'):
        op.create_backward_kernel(target='cpu')
