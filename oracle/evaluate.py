"""CPU oracle — numpy evaluation of stencil assignments.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this package; the
product (``pystencils_autodiff_b200``) never does and has no CPU fallback.

PARITY STATUS: **numerically unpinned at the pystencils boundary.**  The reference evaluates its kernels with
the third-party package pystencils (``pystencils>=0.2.8``, unpinned: /root/reference/setup.cfg:35), which is
neither vendored in /root/reference nor installable here, and the reference's tests hold no numeric golden
vectors for this path.  This file restates pystencils' kernel semantics as used by the
reference and is pinned by (i) the reference's symbolic known answers (tests/test_autodiff.py:21,46;
README.rst:66-68,85-86) via ``tests/test_symbolic.py``, (ii) the reference's own ``_autodiff.py`` executed on
top of our front end (``tests/golden/make_reference_golden.py``), and (iii) gradient checks in the style of
tests/test_tfmad.py:186-231.

Semantics restated:

* ``boundary_handling=None``: iterate ``gl <= c_k < N_k - gl`` with ``gl = max |offset|`` over all accesses of the
  kernel (pystencils ``create_kernel(..., ghost_layers=None)``, call sites _autodiff.py:486-489,501-505);
  outputs are zero elsewhere (``torch.zeros`` at backends/_torch_native.py:64,108).
* ``'zeros'``: iterate every cell; every relative read outside the array is 0
  (transformations.py:12-36; ``ghost_layers=0`` at _autodiff.py:484,499,517,534).
* Arithmetic is carried out in float64 and rounded to the output dtype on store (pystencils'
  ``data_type='double'`` default); pass ``compute_dtype=np.float32`` to mimic ``data_type='float32'``.
"""
import numpy as np
import sympy as sp

from ._model import (as_collection, conditional_accesses_in, field_dtype, ghost_width, index_tail as _index_tail, is_access,
                     offsets_of, spatial_shape)

__all__ = ['evaluate', 'evaluate_literal', 'evaluate_loops', 'forward_backward']


def _mode(boundary_handling):
    v = getattr(boundary_handling, 'value', boundary_handling)
    if v is None:
        return 'none'
    if str(v) == 'zeros':
        return 'zeros'
    raise NotImplementedError(boundary_handling)


def _accesses(ac):
    return ac.reads(), ac.writes()


def _spatial_shape(ac, arrays):
    return spatial_shape(as_collection(ac), arrays)


def evaluate(assignments, arrays, boundary_handling=None, scalars=None, ghost_layers=None,
             compute_dtype=np.float64):
    """Evaluate one kernel.  ``arrays``: name -> ndarray for every field read (and, for ``+=`` forms, written).

    Returns ``{output field name: ndarray}`` (fresh arrays, zero where the kernel does not write).
    """
    ac = as_collection(assignments)
    mode = _mode(boundary_handling)
    scalars = dict(scalars or {})
    reads, writes = _accesses(ac)
    shape = _spatial_shape(ac, arrays)
    ndim = len(shape)
    if mode == 'zeros':
        gl = 0
    elif ghost_layers is not None:
        gl = int(ghost_layers)
    else:
        gl = max([ghost_width(a) for a in reads + writes] + [0])
    pad = max([ghost_width(a) for a in reads] + [0])

    region_lo = [gl] * ndim
    region_hi = [n - gl for n in shape]
    if any(h <= l for l, h in zip(region_lo, region_hi)):
        region_hi = region_lo  # empty

    def read_view(a):
        arr = np.asarray(arrays[a.field.name])
        tail = _index_tail(a)
        if mode == 'zeros':
            padded = np.pad(arr, [(pad, pad)] * ndim + [(0, 0)] * (arr.ndim - ndim))
            sl = tuple(slice(pad + o + l, pad + o + h) for o, l, h in zip(offsets_of(a), region_lo, region_hi))
            v = padded[sl]
        else:
            sl = tuple(slice(o + l, o + h) for o, l, h in zip(offsets_of(a), region_lo, region_hi))
            v = arr[sl]
        if tail:
            v = v[(Ellipsis,) + tail]
        return v.astype(compute_dtype)

    env = {a: read_view(a) for a in reads}
    for s in ac.free_symbols:
        if not is_access(s):
            if s.name not in scalars:
                raise KeyError('missing scalar %s' % s.name)
            env[s] = compute_dtype(scalars[s.name])

    def ev(expr):
        syms = sorted(expr.free_symbols, key=str)
        fn = sp.lambdify(syms, expr, modules='numpy')
        with np.errstate(all='ignore'):
            res = fn(*[env[s] for s in syms])
        region_shape = tuple(h - l for l, h in zip(region_lo, region_hi))
        return np.broadcast_to(np.asarray(res, dtype=compute_dtype), region_shape)

    for a in ac.subexpressions:
        env[a.lhs] = ev(a.rhs)

    out = {}
    for a in ac.main_assignments:
        f = a.lhs.field
        if f.name not in out:
            full_shape = shape + tuple(int(s) for s in f.index_shape)
            out[f.name] = np.zeros(full_shape, dtype=field_dtype(f))
        val = ev(a.rhs)
        sl = tuple(slice(l + o, h + o) for o, l, h in zip(offsets_of(a.lhs), region_lo, region_hi))
        tail = _index_tail(a.lhs)
        out[f.name][sl + tail if tail else sl] = val.astype(field_dtype(f))
    return out


def evaluate_literal(assignments, arrays, scalars=None, compute_dtype=np.float64):
    """Evaluate a collection in the reference's *symbolic* ``'zeros'`` form (``ConditionalFieldAccess`` nodes,
    transformations.py:26-30) literally: every cell, index grids, condition → 0.  Independent of the padding trick
    in :func:`evaluate`; used to cross-check it."""
    ac = as_collection(assignments)
    scalars = dict(scalars or {})
    shape = _spatial_shape(ac, arrays)
    ndim = len(shape)
    grids = np.indices(shape)
    ctr = {sp.Symbol('ctr_%d' % k, integer=True): grids[k] for k in range(ndim)}

    def access_value(a):
        arr = np.asarray(arrays[a.field.name])
        idx = tuple(np.clip(grids[k] + offsets_of(a)[k], 0, shape[k] - 1) for k in range(ndim))
        v = arr[idx + _index_tail(a)] if a.index else arr[idx]
        return v.astype(compute_dtype)

    env = dict(ctr)
    for s in set(ac.free_symbols) | set(_accesses(ac)[0]):
        if is_access(s):
            env[s] = access_value(s)
        elif s not in ctr:
            env[s] = compute_dtype(scalars[s.name])

    def ev(expr):
        repl = {}
        for c in conditional_accesses_in(expr):
            cond_syms = sorted(c.outofbounds_condition.free_symbols, key=str)
            cond = sp.lambdify(cond_syms, c.outofbounds_condition, modules='numpy')(*[env[s] for s in cond_syms])
            d = sp.Dummy()
            env[d] = np.where(cond, compute_dtype(float(c.outofbounds_value)), ev(c.access))
            repl[c] = d
        expr = expr.xreplace(repl)
        syms = sorted(expr.free_symbols, key=str)
        with np.errstate(all='ignore'):
            res = sp.lambdify(syms, expr, modules='numpy')(*[env[s] for s in syms])
        return np.broadcast_to(np.asarray(res, dtype=compute_dtype), shape)

    for a in ac.subexpressions:
        env[a.lhs] = ev(a.rhs)
    out = {}
    for a in ac.main_assignments:
        f = a.lhs.field
        if f.name not in out:
            out[f.name] = np.zeros(shape + tuple(int(s) for s in f.index_shape), dtype=field_dtype(f))
        tail = _index_tail(a.lhs)
        out[f.name][(Ellipsis,) + tail if tail else Ellipsis] = ev(a.rhs).astype(field_dtype(f))
    return out


def evaluate_loops(assignments, arrays, boundary_handling=None, scalars=None):
    """Plain Python loop nest, one cell at a time (tiny shapes only) — the most literal restatement of the
    generated kernels' per-cell body (SURVEY.md Appendix C sketch)."""
    import itertools
    import math
    ac = as_collection(assignments)
    mode = _mode(boundary_handling)
    scalars = dict(scalars or {})
    reads, writes = _accesses(ac)
    shape = _spatial_shape(ac, arrays)
    ndim = len(shape)
    gl = 0 if mode == 'zeros' else max([ghost_width(a) for a in reads + writes] + [0])
    out = {}
    for a in ac.main_assignments:
        f = a.lhs.field
        out.setdefault(f.name, np.zeros(shape + tuple(int(s) for s in f.index_shape), dtype=field_dtype(f)))
    all_syms = sorted(set(ac.free_symbols) | set(reads), key=str)
    sub_fns = [(a.lhs, sp.lambdify(sorted(a.rhs.free_symbols, key=str), a.rhs, modules='math'),
                sorted(a.rhs.free_symbols, key=str)) for a in ac.subexpressions]
    main_fns = [(a.lhs, sp.lambdify(sorted(a.rhs.free_symbols, key=str), a.rhs, modules='math'),
                 sorted(a.rhs.free_symbols, key=str)) for a in ac.main_assignments]
    for c in itertools.product(*[range(gl, n - gl) for n in shape]):
        env = {}
        for s in all_syms:
            if is_access(s):
                idx = tuple(ci + o for ci, o in zip(c, offsets_of(s)))
                if all(0 <= i < n for i, n in zip(idx, shape)):
                    env[s] = float(arrays[s.field.name][idx + _index_tail(s)])
                else:
                    assert mode == 'zeros'
                    env[s] = 0.0
            else:
                env[s] = float(scalars[s.name])
        for lhs, fn, syms in sub_fns:
            env[lhs] = _safe(fn, [env[s] for s in syms], math)
        for lhs, fn, syms in main_fns:
            idx = tuple(ci + o for ci, o in zip(c, offsets_of(lhs)))
            out[lhs.field.name][idx + _index_tail(lhs)] = _safe(fn, [env[s] for s in syms], math)
    return out


def _safe(fn, args, math):
    try:
        return fn(*args)
    except (ValueError, ZeroDivisionError, OverflowError):
        return float('nan')


def forward_backward(op, inputs, grads, scalars=None, compute_dtype=np.float64):
    """Run an ``AutoDiffOp``'s forward and adjoint kernels on numpy arrays the way the reference's autograd
    Function marshals them (backends/_torch_native.py:43-118): backward sees the upstream gradients under the
    ``diff<out>`` names plus every saved forward tensor.

    ``inputs``: name -> array for the forward input fields; ``grads``: name of forward output -> upstream grad.
    Returns ``(outputs, input_grads)`` as dicts keyed by field name.
    """
    bh = op.boundary_handling
    outs = evaluate(op.forward_assignments, inputs, bh, scalars, compute_dtype=compute_dtype)
    env = dict(inputs)
    env.update(outs)
    grad_fields = [f for f in op.backward_input_fields if f not in op.forward_input_fields]
    for f in grad_fields:
        fwd_name = f.corresponding_forward_field.name if hasattr(f, 'corresponding_forward_field') else f.name[4:]
        env[f.name] = grads[fwd_name]
    # ``+=`` forms (time-constant fields) read their own output: start from zeros like torch.zeros does
    for f in op.backward_output_fields:
        if f.name not in env:
            shape = _spatial_shape(op.backward_assignments, env)
            env[f.name] = np.zeros(shape + tuple(int(s) for s in f.index_shape), dtype=field_dtype(f))
    din = evaluate(op.backward_assignments, env, bh, scalars, compute_dtype=compute_dtype)
    return outs, din
