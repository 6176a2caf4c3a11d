"""Generate golden fixtures by executing the REFERENCE's own symbolic layer.

Runs only in the build container (needs /root/reference).  pystencils — the reference's third-party dependency — is
not installable, so the reference's ``_autodiff.py`` / ``_adjoint_field.py`` / ``transformations.py`` are imported
*unmodified from where they lie* with ``pystencils`` resolved to a shim that maps the handful of names they use
onto this repo's front end (``Field``, ``Assignment``, ``AssignmentCollection``, CSE, ``ConditionalFieldAccess``).
What this pins: the TF-MAD / transposed differentiation rules, field orderings and the 'zeros' boundary transform of
the reference itself — not pystencils' code generation, which stays unpinned (see oracle/evaluate.py).

Outputs (committed): tests/golden/reference_symbolic.json, tests/golden/reference_numeric.npz

    python tests/golden/make_reference_golden.py
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference/src/pystencils_autodiff'

import pystencils_autodiff_b200 as ours  # noqa: E402
from pystencils_autodiff_b200 import assignment as A, field as F, transformations as T  # noqa: E402


def install_shim():
    ps = types.ModuleType('pystencils')
    class ShimField(F.Field):
        """pystencils' constructor derives the index dimensions from ``len(shape) - len(layout)`` (the reference's
        AdjointField relies on it for vector fields); this front end sets them in its factory functions instead."""

        def __init__(self, name, field_type, dtype, layout, shape, strides=None):
            super().__init__(name, field_type, dtype, layout, shape, strides)
            self._index_dimensions = len(self.shape) - len(self._layout)
    ps.Field, ps.FieldType, ps.fields = ShimField, F.FieldType, F.fields
    ps.Assignment, ps.AssignmentCollection = A.Assignment, A.AssignmentCollection
    ps.x_vector = F.x_vector
    sub = {}
    for name in ['cache', 'interpolation_astnodes', 'math_optimizations', 'data_types', 'simp', 'astnodes', 'field']:
        m = types.ModuleType('pystencils.' + name)
        sub[name] = m
        setattr(ps, name, m)
        sys.modules['pystencils.' + name] = m
    sub['cache'].disk_cache_no_fallback = lambda fn: fn
    sub['interpolation_astnodes'].InterpolatorAccess = type('InterpolatorAccess', (sp.Function,), {})
    sub['math_optimizations'].ReplaceOptim = lambda *a, **k: None
    sub['math_optimizations'].optimize_assignments = lambda assignments, optims: list(assignments)
    sub['data_types'].cast_func = type('cast_func', (sp.Function,), {})
    sub['simp'].sympy_cse_on_assignment_list = A.sympy_cse_on_assignment_list
    sub['simp'].sympy_cse = A.sympy_cse
    sub['astnodes'].ConditionalFieldAccess = T.ConditionalFieldAccess
    sub['astnodes'].FieldShapeSymbol = type('FieldShapeSymbol', (sp.Symbol,), {})
    sub['astnodes'].FieldStrideSymbol = type('FieldStrideSymbol', (sp.Symbol,), {})
    sub['field'].Field = F.Field
    ps.create_kernel = None
    sys.modules['pystencils'] = ps

    pkg = types.ModuleType('pystencils_autodiff')
    pkg.__path__ = [REF]
    sys.modules['pystencils_autodiff'] = pkg
    lf = types.ModuleType('pystencils_autodiff._layout_fixer')
    sys.modules['pystencils_autodiff._layout_fixer'] = lf
    pkg._layout_fixer = lf
    be = types.ModuleType('pystencils_autodiff.backends')
    be.AVAILABLE_BACKENDS = ['tensorflow', 'torch', 'tensorflow_native', 'torch_native']
    sys.modules['pystencils_autodiff.backends'] = be

    def load(modname, filename):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, filename))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    adj = load('pystencils_autodiff._adjoint_field', '_adjoint_field.py')
    pkg.AdjointField = adj.AdjointField
    tr = load('pystencils_autodiff.transformations', 'transformations.py')
    ad = load('pystencils_autodiff._autodiff', '_autodiff.py')
    return ad, tr


def cases():
    """name -> (assignment factory, kwargs for AutoDiffOp, value range for numeric inputs)"""
    from pystencils_autodiff_b200.configs import (diffusion2d_op, heat3d_op, readme_op, stencil27_op, tv_gradient_op)
    c = {}
    c['readme'] = (lambda: readme_op((6, 7), 'float64').forward_assignments, {}, (0.5, 1.5))
    c['diffusion2d'] = (lambda: diffusion2d_op((7, 9), 'float64').forward_assignments, {}, (-1, 1))
    c['heat3d'] = (lambda: heat3d_op((5, 6, 7), 'float64').forward_assignments, {}, (-1, 1))
    c['stencil27'] = (lambda: stencil27_op((5, 6, 7), 'float64').forward_assignments, {}, (-1, 1))
    c['tvgrad'] = (lambda: tv_gradient_op((2, 7, 9), 'float64').forward_assignments, {}, (0, 1))

    def fd_stencil():  # /root/reference/tests/test_tfmad.py:191-200
        a, b, out = ours.fields('a, b, out: float64[5,7]')
        cont = 2 * ours.fd.Diff(a, 0) - 1.5 * ours.fd.Diff(a, 1) - ours.fd.Diff(b, 0) + 3 * ours.fd.Diff(b, 1)
        return ours.AssignmentCollection([ours.Assignment(out.center(), ours.fd.Discretization2ndOrder(dx=1)(cont)
                                                          + 1.2 * a.center())], [])
    c['fd_stencil'] = (fd_stencil, {}, (-1, 1))

    def three_outputs():  # /root/reference/tests/test_tfmad.py:242-248
        a, b, out1, out2, out3 = ours.fields('a, b, out1, out2, out3: float64[9,8]')
        return ours.AssignmentCollection({out1.center: a.center + b.center, out2.center: a.center - b.center,
                                          out3.center: sp.exp(b[-1, 0])})
    c['three_outputs'] = (three_outputs, {}, (-1, 1))

    def quirk():  # field-dependent coefficient at an offset
        x, y, z = ours.fields('x, y, z: float64[6,5]')
        return ours.AssignmentCollection({z.center: x[1, 0] * y[0, 0] + sp.sin(x[0, -1])})
    c['unshifted_coefficient'] = (quirk, {}, (0.5, 1.5))

    def with_subexpr():
        x, y, z = ours.fields('x, y, z: float64[6,5]')
        t0 = sp.Symbol('t0')
        return ours.AssignmentCollection([ours.Assignment(z.center, t0 * x[0, 1] + t0 ** 2)],
                                         [ours.Assignment(t0, y[-1, 0] * 2 + x[0, 0])])
    c['with_subexpression'] = (with_subexpr, {}, (0.5, 1.5))
    c['time_constant'] = (lambda: diffusion2d_op((7, 9), 'float64').forward_assignments, 'time_constant', (-1, 1))
    c['constant_field'] = (quirk, 'constant_y', (0.5, 1.5))
    c['transposed_pointwise'] = (lambda: readme_op((6, 7), 'float64').forward_assignments,
                                 {'diff_mode': 'transposed'}, (0.5, 1.5))
    # rows of a multiple of 16 bytes: on the GPU these take the march (TMA-staged) kernels, not the generic fallback
    c['diffusion2d_aligned'] = (lambda: diffusion2d_op((20, 36), 'float64').forward_assignments, {}, (-1, 1))
    c['heat3d_aligned'] = (lambda: heat3d_op((6, 10, 36), 'float64').forward_assignments, {}, (-1, 1))
    c['stencil27_aligned'] = (lambda: stencil27_op((5, 9, 36), 'float64').forward_assignments, {}, (-1, 1))
    c['tvgrad_aligned'] = (lambda: tv_gradient_op((2, 12, 36), 'float64').forward_assignments, {}, (0, 1))
    # random multi-field stencils (tests/stencil_fuzz.py: products, exp, sqrt, sin; offsets up to +-4), float64 seeds only
    for seed in RANDOM_SEEDS:
        c['random_%d' % seed] = (lambda s=seed: random_case(s), {}, (0.5, 1.5))
    return c


def symbolic_only_cases():
    """Cases compared as printed assignments only (no numeric vectors): index-dimension outputs and symbolic shapes.
    name -> (assignment factory, kwargs for AutoDiffOp)"""
    c = {}

    def curl():  # /root/reference/tests/test_tfmad.py:341-401 (get_curl + test_tfmad_two_outputs): scalar -> vector field
        u = ours.Field.create_fixed_size('curl_input', (20, 30), index_dimensions=0)
        cf = ours.Field.create_fixed_size('curl', (20, 30, 2), index_dimensions=1)
        disc = ours.fd.Discretization2ndOrder(dx=1)
        return ours.AssignmentCollection([ours.Assignment(cf.center(0), disc(ours.fd.Diff(u, 0))),
                                          ours.Assignment(cf.center(1), disc(ours.fd.Diff(u, 1)))], [])
    c['curl_vector_output'] = (curl, {'diff_mode': 'transposed-forward'})

    def one_stencil():  # /root/reference/tests/test_tfmad.py:12-28, symbolic shapes
        f, out = ours.fields('f, out: double[2D]')
        cont = ours.fd.Diff(f, 0) - ours.fd.Diff(f, 1)
        return ours.AssignmentCollection([ours.Assignment(out.center(), ours.fd.Discretization2ndOrder(dx=1)(cont))], [])
    c['tfmad_stencil_2d_symbolic_shape'] = (one_stencil, {'diff_mode': 'transposed-forward'})

    def two_stencils():  # /root/reference/tests/test_tfmad.py:31-53
        a, b, out = ours.fields('a, b, out: double[2D]')
        cont = ours.fd.Diff(a, 0) - ours.fd.Diff(a, 1) - ours.fd.Diff(b, 0) + ours.fd.Diff(b, 1)
        return ours.AssignmentCollection([ours.Assignment(out.center(), ours.fd.Discretization2ndOrder(dx=1)(cont))], [])
    c['tfmad_two_stencils_2d_symbolic_shape'] = (two_stencils, {'diff_mode': 'transposed-forward'})
    return c


RANDOM_SEEDS = (0, 4, 5, 8, 13, 21, 24, 31)


def random_case(seed):
    sys.path.insert(0, os.path.dirname(HERE))
    from stencil_fuzz import random_stencil
    _, _, shape, dtype = random_stencil(seed)
    assert dtype == 'float64', 'golden vectors are float64: pick another seed'
    fa, _, _, _ = random_stencil(seed, shape=(11, 12) if len(shape) == 2 else (6, 7, 12))
    return fa


def resolve_kwargs(spec, fa):
    if spec == 'time_constant':
        f = sorted(fa.free_fields, key=str)[0]
        return {'time_constant_fields': [f]}
    if spec == 'constant_y':
        f = [f for f in fa.free_fields if f.name == 'y']
        return {'constant_fields': list(f)}
    return dict(spec)


def main():
    # The reference's 'transposed' mode orders fields and assignments by SET iteration (_autodiff.py:430 ``list(read_fields)``),
    # i.e. by string hashes: pin the hash seed so that regenerating the fixtures is reproducible.  (The consumer,
    # tests/test_symbolic.py, compares that case order-insensitively.)
    if os.environ.get('PYTHONHASHSEED') != '0':
        os.environ['PYTHONHASHSEED'] = '0'
        os.execv(sys.executable, [sys.executable] + sys.argv)
    from oracle import evaluate      # here, not at import time: the case list is also read by scripts/precompile_tests.py
    ad, tr = install_shim()
    sym, num = {}, {}
    rng = np.random.default_rng(2026)
    for name, (factory, spec, (lo, hi)) in cases().items():
        fa = factory()
        kw = resolve_kwargs(spec, fa)
        ref_op = ad.AutoDiffOp(fa, **kw)
        entry = {
            'forward': str(ref_op.forward_assignments),
            'backward': str(ref_op.backward_assignments),
            'forward_input_fields': [f.name for f in ref_op.forward_input_fields],
            'forward_output_fields': [f.name for f in ref_op.forward_output_fields],
            'backward_input_fields': sorted(f.name for f in ref_op.backward_input_fields),
            'backward_output_fields': sorted(f.name for f in ref_op.backward_output_fields),
        }
        # the reference's 'zeros' transform on its own backward assignments
        bh = tr.add_fixed_constant_boundary_handling(ref_op.backward_assignments)
        entry['backward_zeros'] = str(bh)
        sym[name] = entry
        # numeric golden vectors: the reference-produced assignments evaluated by the oracle on seeded inputs
        shape = tuple(int(s) for s in sorted(fa.free_fields, key=str)[0].spatial_shape)
        for mode in (None, 'zeros'):
            arrays = {f.name: rng.uniform(lo, hi, size=shape) for f in ref_op.forward_input_fields}
            outs = evaluate(ref_op.forward_assignments, arrays, mode)
            env = dict(arrays)
            env.update(outs)
            for f in ref_op.backward_input_fields:
                if f.name not in env:
                    env[f.name] = rng.normal(size=shape)
            for f in ref_op.backward_output_fields:
                env.setdefault(f.name, np.zeros(shape))
            grads = evaluate(ref_op.backward_assignments, env, mode)
            tag = '%s/%s/' % (name, 'none' if mode is None else 'zeros')
            for k, v in env.items():
                if k not in grads or k in [f.name for f in ref_op.backward_input_fields]:
                    num[tag + 'in/' + k] = v
            for k, v in outs.items():
                num[tag + 'out/' + k] = v
            for k, v in grads.items():
                num[tag + 'grad/' + k] = v
    for name, (factory, kw) in symbolic_only_cases().items():
        ref_op = ad.AutoDiffOp(factory(), **kw)
        entry = {
            'forward': str(ref_op.forward_assignments),
            'backward': str(ref_op.backward_assignments),
            'forward_input_fields': [f.name for f in ref_op.forward_input_fields],
            'forward_output_fields': [f.name for f in ref_op.forward_output_fields],
            'backward_input_fields': sorted(f.name for f in ref_op.backward_input_fields),
            'backward_output_fields': sorted(f.name for f in ref_op.backward_output_fields),
            'symbolic_only': True,
        }
        try:
            entry['backward_zeros'] = str(tr.add_fixed_constant_boundary_handling(ref_op.backward_assignments))
        except Exception as exc:          # symbolic shapes: recorded, the consumer expects the same failure
            entry['backward_zeros_error'] = type(exc).__name__
        sym[name] = entry
    with open(os.path.join(HERE, 'reference_symbolic.json'), 'w') as fh:
        json.dump(sym, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, 'reference_numeric.npz'), **num)
    print('wrote %d symbolic cases, %d arrays' % (len(sym), len(num)))


if __name__ == '__main__':
    main()
