"""The oracle checked against an evaluator-independent second opinion, and its independence from the product.

``oracle/handwritten.py`` writes the five BASELINE.json stencils and their TF-MAD adjoints out with explicit numpy shifts
(partial derivatives of the non-linear ones by complex-step differentiation): no sympy, no ``oracle.evaluate``, no product
code.  Agreement here pins (i) the evaluator's offset / boundary / iteration-space semantics, (ii) sympy's derivatives as
used by the differentiation rules, (iii) the un-shifted-coefficient rule of _autodiff.py:104-109 on a non-linear stencil."""
import ast
import os

import numpy as np
import pytest
import sympy as sp

from oracle import evaluate
from oracle.cgen import compile_c
from oracle.handwritten import HANDWRITTEN, shifted
from pystencils_autodiff_b200.configs import make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = {'c1': (9, 11), 'c2': (12, 20), 'c3': (6, 9, 12), 'c4': (5, 8, 12), 'c5': (3, 10, 12)}


def _arrays(op, shape, seed):
    rng = np.random.default_rng(seed)
    return {f.name: rng.uniform(0.5, 1.5, size=shape) for f in sorted(set(op.forward_fields) | set(op.backward_fields), key=str)}


@pytest.mark.parametrize('bh', ['zeros', None])
@pytest.mark.parametrize('name', sorted(SHAPES))
def test_evaluator_matches_handwritten_restatement(name, bh):
    op = make_config(name, shape=SHAPES[name], dtype='float64', boundary_handling=bh)
    arrs = _arrays(op, SHAPES[name], 11)
    fwd, bwd = HANDWRITTEN[name]
    for assigns, hand in ((op.forward_assignments, fwd), (op.backward_assignments, bwd)):
        got, want = evaluate(assigns, arrs, bh), hand(arrs, bh)
        assert set(got) == set(want)
        for k in got:
            assert np.abs(got[k] - want[k]).max() <= 1e-13 * max(1.0, np.abs(want[k]).max()), (name, bh, k)
            assert np.abs(want[k]).max() > 0.1                                   # not a comparison of zeros


@pytest.mark.parametrize('name', ['c2', 'c3', 'c4', 'c5'])
def test_c_restatement_matches_handwritten(name):
    """The C/OpenMP restatement (the CPU baseline bench.py times and the full-size checker) against the same hand-written
    formulas, in the field dtype of the BASELINE config."""
    dt = 'float64' if name == 'c4' else 'float32'
    shape = SHAPES[name]
    op = make_config(name, shape=shape, dtype=dt, boundary_handling='zeros')
    arrs = {k: v.astype(dt) for k, v in _arrays(op, shape, 5).items()}
    fwd, bwd = HANDWRITTEN[name]
    tol = 1e-13 if dt == 'float64' else 2e-6
    for assigns, hand, tag in ((op.forward_assignments, fwd, 'forward'), (op.backward_assignments, bwd, 'backward')):
        k = compile_c(assigns, 'zeros', '%s_%s_hw' % (op.op_name, tag), 'strict')
        want = hand(arrs, 'zeros')
        bufs = {n: (arrs[n].copy() if n not in want else np.zeros(shape, dtype=dt)) for n in k.field_names}
        k(**bufs)
        for n in want:
            assert np.abs(bufs[n] - want[n]).max() <= tol * max(1.0, np.abs(want[n]).max()), (name, tag, n)


def test_exact_adjoint_mode_matches_the_transposed_jacobian_by_hand():
    op = make_config('c5', shape=SHAPES['c5'], dtype='float64', boundary_handling='zeros', adjoint_mode='exact')
    arrs = _arrays(op, SHAPES['c5'], 3)
    got = evaluate(op.backward_assignments, arrs, 'zeros')
    want = HANDWRITTEN['c5'][1](arrs, 'zeros', exact=True)
    assert np.abs(got['diffu'] - want['diffu']).max() <= 1e-12 * np.abs(want['diffu']).max()
    # ... and that one IS the adjoint: <J v, w> == <v, J^T w> by central differences of the hand-written forward
    rng = np.random.default_rng(9)
    v, w = rng.normal(size=SHAPES['c5']), arrs['diffg']
    fwd = HANDWRITTEN['c5'][0]
    h = 1e-6
    jv = (fwd(dict(arrs, u=arrs['u'] + h * v), 'zeros')['g'] - fwd(dict(arrs, u=arrs['u'] - h * v), 'zeros')['g']) / (2 * h)
    assert abs(np.vdot(jv, w) - np.vdot(v, want['diffu'])) <= 1e-6 * abs(np.vdot(jv, w))


def test_shifted_is_the_zero_padded_shift():
    a = np.arange(12.0).reshape(3, 4)
    assert shifted(a, (1, 0))[0].tolist() == a[1].tolist() and shifted(a, (1, 0))[2].tolist() == [0, 0, 0, 0]
    assert shifted(a, (0, -1))[:, 0].tolist() == [0, 0, 0] and shifted(a, (0, -1))[:, 1].tolist() == a[:, 0].tolist()
    assert not shifted(a, (5, 0)).any()


def test_oracle_does_not_import_the_product():
    """Checker and product share no code: no module under oracle/ imports pystencils_autodiff_b200."""
    for fn in sorted(os.listdir(os.path.join(ROOT, 'oracle'))):
        if not fn.endswith('.py'):
            continue
        with open(os.path.join(ROOT, 'oracle', fn)) as fh:
            tree = ast.parse(fh.read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.level == 0:
                mods = [node.module or '']
            assert not any(m.split('.')[0] == 'pystencils_autodiff_b200' for m in mods), (fn, mods)


def test_oracle_evaluates_foreign_objects_through_the_pystencils_protocol():
    """Stand-ins that share nothing with the product's front end: the oracle only reads ``.field/.offsets/.index`` of an
    access, ``.name/.dtype/.spatial_dimensions/.index_shape`` of a field, ``.lhs/.rhs`` of an assignment."""
    class Dt:
        numpy_dtype = np.dtype(np.float64)

    class F:
        def __init__(self, name):
            self.name, self.dtype, self.spatial_dimensions, self.index_shape, self.has_fixed_shape = name, Dt, 2, (), False

        def __str__(self):
            return self.name

    class Acc(sp.Symbol):
        def __new__(cls, field, offsets):
            obj = sp.Symbol.__new__(cls, '%s_%s' % (field.name, '_'.join(str(o).replace('-', 'm') for o in offsets)))
            obj.field, obj.offsets, obj.index = field, tuple(offsets), ()
            return obj

    class Asg:
        def __init__(self, lhs, rhs):
            self.lhs, self.rhs = lhs, rhs

    a, out = F('a'), F('out')
    asg = [Asg(Acc(out, (0, 0)), 2 * Acc(a, (1, 0)) - Acc(a, (0, -1)) * sp.Symbol('k'))]
    A = np.random.default_rng(0).normal(size=(6, 7))
    got = evaluate(asg, {'a': A}, 'zeros', scalars={'k': 3.0})['out']
    want = 2 * shifted(A, (1, 0)) - 3.0 * shifted(A, (0, -1))
    assert np.array_equal(got, want)
    got = evaluate(asg, {'a': A}, None, scalars={'k': 3.0})['out']
    assert np.array_equal(got[1:-1, 1:-1], want[1:-1, 1:-1]) and not got[0].any() and not got[:, -1].any()
