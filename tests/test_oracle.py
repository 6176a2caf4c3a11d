"""The oracle against itself, against the golden vectors and against the known properties of the path."""
import os

import numpy as np
import pytest

import pystencils_autodiff_b200 as ps
from oracle import evaluate, evaluate_literal, evaluate_loops, forward_backward
from oracle.cgen import compile_c
from pystencils_autodiff_b200.configs import make_config

HERE = os.path.dirname(__file__)
SMALL = {'c1': (6, 7), 'c2': (7, 9), 'c3': (5, 6, 7), 'c4': (5, 6, 7), 'c5': (2, 7, 9)}
RANGE = {'c1': (0.5, 1.5), 'c2': (-1, 1), 'c3': (-1, 1), 'c4': (-1, 1), 'c5': (0, 1)}


def _inputs(op, shape, lo, hi, seed=0):
    rng = np.random.default_rng(seed)
    ins = {f.name: rng.uniform(lo, hi, size=shape).astype(f.dtype.numpy_dtype) for f in op.forward_input_fields}
    grads = {f.name: rng.normal(size=shape).astype(f.dtype.numpy_dtype) for f in op.forward_output_fields}
    return ins, grads


@pytest.mark.parametrize('bh', [None, 'zeros'])
@pytest.mark.parametrize('name', sorted(SMALL))
def test_vectorised_vs_loops_vs_c(name, bh):
    op = make_config(name, shape=SMALL[name], dtype='float64', boundary_handling=bh)
    ins, grads = _inputs(op, SMALL[name], *RANGE[name])
    a = evaluate(op.forward_assignments, ins, bh)
    b = evaluate_loops(op.forward_assignments, ins, bh)
    k = compile_c(op.forward_assignments, bh, 'fwd_' + name, flavour='strict')
    for f in op.forward_output_fields:
        c = np.zeros(SMALL[name])
        k(**ins, **{f.name: c})
        np.testing.assert_allclose(a[f.name], b[f.name], rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(a[f.name], c, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize('name', ['c2', 'c3', 'c5'])
def test_zero_padding_equals_literal_conditional_form(name):
    """transformations.py:12-36 evaluated literally (index grids + conditions) == padded evaluation."""
    op = make_config(name, shape=SMALL[name], dtype='float64', boundary_handling='zeros')
    ins, grads = _inputs(op, SMALL[name], *RANGE[name])
    fwd_sym, bwd_sym = op.symbolic_boundary_handled_assignments()
    a = evaluate(op.forward_assignments, ins, 'zeros')
    b = evaluate_literal(fwd_sym, ins)
    for k in a:
        np.testing.assert_allclose(a[k], b[k], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('gl', [1, 2, 3])
def test_fixed_constant_bh_against_bruteforce(gl):
    """tests/test_fixed_constant_bh.py:22-48 with real assertions: averaging stencil with gl ghost layers."""
    import itertools
    import sympy as sp
    x, y = ps.fields('x, y: float64[12,13]')
    offsets = list(itertools.product(range(gl + 1), repeat=2))
    asg = ps.AssignmentCollection({y.center: sp.Add(*[x[o] for o in offsets]) / len(offsets)})
    noise = np.random.default_rng(gl).random((12, 13))
    plain = evaluate(asg, dict(x=noise), None)['y']
    bh = evaluate_literal(ps.add_fixed_constant_boundary_handling(asg), dict(x=noise))['y']
    padded = np.pad(noise, gl)
    brute = sum(padded[gl + o[0]:gl + o[0] + 12, gl + o[1]:gl + o[1] + 13] for o in offsets) / len(offsets)
    np.testing.assert_allclose(bh, brute, rtol=1e-13)
    np.testing.assert_allclose(plain[gl:-gl, gl:-gl], brute[gl:-gl, gl:-gl], rtol=1e-13)
    assert np.all(plain[:gl] == 0) and np.all(plain[:, -gl:] == 0)


def _dense_operators(op, shape, bh):
    n = int(np.prod(shape))
    names = [f.name for f in op.forward_input_fields]
    J = np.zeros((n, n * len(names)))
    zeros = {nm: np.zeros(shape) for nm in names}
    for j, nm in enumerate(names):
        for i in range(n):
            e = np.zeros(n)
            e[i] = 1
            ins = dict(zeros)
            ins[nm] = e.reshape(shape)
            J[:, j * n + i] = evaluate(op.forward_assignments, ins, bh)[op.forward_output_fields[0].name].ravel()
    Jt = np.zeros((n * len(names), n))
    for i in range(n):
        e = np.zeros(n)
        e[i] = 1
        _, d = forward_backward(op, zeros, {op.forward_output_fields[0].name: e.reshape(shape)})
        Jt[:, i] = np.concatenate([d['diff' + nm].ravel() for nm in names])
    return J, Jt


def test_backward_is_exact_transpose_in_zeros_mode_only():
    """with 'zeros' the adjoint of a linear stencil is the exact transpose; with None it
    differs next to the border (max abs diff 1.5 for the stencil of tests/test_tfmad.py:195-200)."""
    a, b, out = ps.fields("a, b, out: float64[5,7]")
    cont = 2 * ps.fd.Diff(a, 0) - 1.5 * ps.fd.Diff(a, 1) - ps.fd.Diff(b, 0) + 3 * ps.fd.Diff(b, 1)
    asg = ps.Assignment(out.center(), ps.fd.Discretization2ndOrder(dx=1)(cont) + 1.2 * a.center())
    J, Jt = _dense_operators(ps.AutoDiffOp([asg], boundary_handling='zeros'), (5, 7), 'zeros')
    assert np.abs(J.T - Jt).max() == 0
    J, Jt = _dense_operators(ps.AutoDiffOp([asg], boundary_handling=None), (5, 7), None)
    assert np.isclose(np.abs(J.T - Jt).max(), 1.5)


def test_unshifted_coefficient_quirk_is_reproduced():
    """for z = x[1,0]*y the reference's rule gives diffx_C = y_C*diffz_W (not y_W*diffz_W)."""
    x, y, z = ps.fields('x, y, z: float64[6,5]')
    op = ps.AutoDiffOp(ps.AssignmentCollection({z.center: x[1, 0] * y[0, 0]}), boundary_handling='zeros')
    rng = np.random.default_rng(0)
    X, Y, G = rng.uniform(0.5, 1.5, (6, 5)), rng.uniform(0.5, 1.5, (6, 5)), rng.normal(size=(6, 5))
    _, d = forward_backward(op, dict(x=X, y=Y), dict(z=G))
    Gw = np.zeros_like(G)
    Gw[1:] = G[:-1]
    np.testing.assert_allclose(d['diffx'], Y * Gw, rtol=1e-14)


def test_golden_vectors():
    """Fixtures generated from the reference's own backward assignments (tests/golden/make_reference_golden.py)."""
    import sys
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    from make_reference_golden import cases, resolve_kwargs
    gold = np.load(os.path.join(HERE, 'golden', 'reference_numeric.npz'))
    checked = 0
    for name, (factory, spec, _) in cases().items():
        fa = factory()
        op = ps.AutoDiffOp(fa, **resolve_kwargs(spec, fa))
        for mode in (None, 'zeros'):
            tag = '%s/%s/' % (name, 'none' if mode is None else 'zeros')
            env = {k[len(tag) + 3:]: gold[k] for k in gold.files if k.startswith(tag + 'in/')}
            outs = evaluate(op.forward_assignments, env, mode)
            for k, v in outs.items():
                np.testing.assert_allclose(v, gold[tag + 'out/' + k], rtol=1e-13, atol=1e-13)
            for f in op.backward_output_fields:   # ``+=`` forms read their own (zero-initialised) output
                env.setdefault(f.name, np.zeros(next(iter(env.values())).shape))
            grads = evaluate(op.backward_assignments, env, mode)
            for k, v in grads.items():
                np.testing.assert_allclose(v, gold[tag + 'grad/' + k], rtol=1e-13, atol=1e-13)
                checked += 1
    assert checked >= 24


def test_c_restatement_against_golden_vectors():
    """oracle/cgen.py — the C/OpenMP restatement of the pystencils CPU loop nest that ``bench.py --impl reference`` times
    — against the committed golden vectors (strict flavour; the fast flavour is the same source with -Ofast)."""
    import sys
    sys.path.insert(0, HERE)
    from golden_util import build_op, golden_arrays, golden_names
    checked = 0
    for name in golden_names():
        if name.startswith('random_') and name not in ('random_0', 'random_4'):
            continue                                   # two random stencils are enough for the C generator
        for mode in (None, 'zeros'):
            op = build_op(name, mode)
            ins, outs, grads = golden_arrays(name, mode)
            shape = next(iter(ins.values())).shape
            for assigns, gold, tag in ((op.forward_assignments, outs, 'f'), (op.backward_assignments, grads, 'b')):
                k = compile_c(assigns, mode, 'gold_%s_%s_%s' % (name, tag, 'z' if mode else 'n'), flavour='strict')
                env = {}
                for n in k.field_names:
                    env[n] = np.ascontiguousarray(ins[n]) if (n in ins and n not in gold) else np.zeros(shape)
                k(**env)
                for key, ref in gold.items():
                    np.testing.assert_allclose(env[key], ref, rtol=1e-12, atol=1e-12, err_msg='%s %s %s' % (name, mode, key))
                    checked += 1
    assert checked >= 60


def _fd_vjp_error(op, shape, seed=0, eps=1e-6, border=0):
    """max |J^T g - backward(g)| with J from central finite differences of the oracle's forward evaluation; ``border``:
    upstream gradients are zero within that many cells of the array border."""
    rng = np.random.default_rng(seed)
    bh = op.boundary_handling
    ins = {f.name: rng.uniform(0.5, 1.5, shape) for f in op.forward_input_fields}
    grads = {f.name: rng.standard_normal(shape) for f in op.forward_output_fields}
    if border:
        for g in grads.values():
            mask = np.zeros(shape, dtype=bool)
            mask[tuple(slice(border, -border) for _ in shape)] = True
            g[~mask] = 0.0
    _, din = forward_backward(op, ins, grads)
    worst = 0.0
    for f in op.forward_input_fields:
        key = 'diff' + f.name
        if key not in din:
            continue
        fd = np.zeros(shape)
        for idx in np.ndindex(*shape):
            plus, minus = {k: v.copy() for k, v in ins.items()}, {k: v.copy() for k, v in ins.items()}
            plus[f.name][idx] += eps
            minus[f.name][idx] -= eps
            op_, om_ = evaluate(op.forward_assignments, plus, bh), evaluate(op.forward_assignments, minus, bh)
            fd[idx] = sum(((op_[o] - om_[o]) * grads[o]).sum() for o in grads) / (2 * eps)
        worst = max(worst, float(np.abs(fd - din[key]).max()))
    return worst


def test_exact_adjoint_mode_passes_the_gradient_check_for_nonlinear_stencils():
    """the reference leaves field-dependent coefficients at the centre cell, which is not
    the transpose of the Jacobian when the read is at an offset.  ``adjoint_mode='exact'`` shifts them: finite-difference
    vector-Jacobian products agree for the TV gradient (C5) and a two-field non-linear stencil with an off-centre write;
    the default mode does not, and for linear stencils / centre reads the two modes emit the same assignments."""
    import sympy as sp
    from pystencils_autodiff_b200 import configs
    shape = (2, 5, 6)
    tv = {m: configs.tv_gradient_op(shape=shape, dtype='float64', boundary_handling='zeros', adjoint_mode=m)
          for m in ('reference', 'exact')}
    assert _fd_vjp_error(tv['exact'], shape) < 1e-6
    assert _fd_vjp_error(tv['reference'], shape) > 1e-2

    x, y, z = ps.fields('x, y, z: float64[5,6]')
    asg = ps.AssignmentCollection({z.center: x[1, 0] * y[0, 0] + sp.sin(x[0, -1]) * y[-1, 1]})
    err = {m: _fd_vjp_error(ps.AutoDiffOp(asg, boundary_handling='zeros', adjoint_mode=m), (5, 6), seed=1)
           for m in ('reference', 'exact')}
    assert err['exact'] < 1e-6 < 1e-2 < err['reference']
    # an off-centre write (interior iteration; upstream gradients away from the border): diff_out is read at -o + l
    X, Y, Z = ps.fields('X, Y, Z: float64[11,12]')
    off = ps.AssignmentCollection({Z[0, 1]: X[1, 0] * Y[0, 0] + 0.5 * X[0, -1]})
    err = {m: _fd_vjp_error(ps.AutoDiffOp(off, boundary_handling=None, adjoint_mode=m), (11, 12), seed=2, border=4)
           for m in ('reference', 'exact')}
    assert err['exact'] < 1e-6 < 1e-2 < err['reference']

    for make, shp in ((configs.heat3d_op, (4, 5, 6)), (configs.stencil27_op, (4, 5, 6)), (configs.readme_op, (5, 6))):
        a = make(shape=shp, adjoint_mode='reference').backward_assignments
        b = make(shape=shp, adjoint_mode='exact').backward_assignments
        assert str(a) == str(b)
    with pytest.raises(ValueError):
        ps.AutoDiffOp(asg, adjoint_mode='almost')
    with pytest.raises(NotImplementedError):
        ps.AutoDiffOp(ps.AssignmentCollection({z.center: x.center * y.center}), diff_mode='transposed', adjoint_mode='exact')
