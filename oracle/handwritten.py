"""Hand-written numpy restatements of the five BASELINE.json stencils and their adjoints.  TEST INFRASTRUCTURE.

An evaluator-independent second opinion ("the numeric evaluator is checked only against itself"): nothing here
goes through sympy, through ``oracle.evaluate`` or through any product code.  Every stencil is written out with explicit
array shifts from its definition in SURVEY.md §8(d) / BASELINE.json, and every adjoint from the reference's TF-MAD rule

    diff_f[c] = sum over read accesses ra = f[o] of  (d rhs / d ra)(c) * diff_out[c - o]          (_autodiff.py:88-109)

(the coefficient stays at cell c — SURVEY Appendix B-1), with the partial derivatives of the non-linear stencils taken by
COMPLEX-STEP differentiation of the hand-written right-hand side (exact to rounding, no symbolic algebra involved).

Boundary modes: ``'zeros'`` — every cell is evaluated, reads outside the array are 0
(transformations.py:12-36), in the backward kernel for forward fields and upstream gradients alike; ``None`` — only cells at
least ``gl = max |offset|`` (of that kernel's own accesses) from every border are evaluated, the rest of the output is 0
(torch.zeros at backends/_torch_native.py:64,108).
"""
import itertools

import numpy as np

__all__ = ['shifted', 'HANDWRITTEN']


def shifted(a, offset):
    """``S[c] = a[c + offset]`` where ``c + offset`` is inside the array, else 0.  ``offset`` has one entry per axis."""
    a = np.asarray(a)
    out = np.zeros_like(a)
    src, dst = [], []
    for n, o in zip(a.shape, offset):
        o = int(o)
        if abs(o) >= n:
            return out
        src.append(slice(max(o, 0), n + min(o, 0)))
        dst.append(slice(max(-o, 0), n - max(o, 0)))
    out[tuple(dst)] = a[tuple(src)]
    return out


def _interior(a, gl):
    """Keep cells at least ``gl`` from every border, zero the rest (``boundary_handling=None``)."""
    if gl == 0:
        return a
    out = np.zeros_like(a)
    if all(n > 2 * gl for n in a.shape):
        sl = tuple(slice(gl, n - gl) for n in a.shape)
        out[sl] = a[sl]
    return out


def _finish(res, gl, boundary, dtype):
    res = {k: (v if boundary == 'zeros' else _interior(v, gl)).astype(dtype) for k, v in res.items()}
    return res


def _neg(o):
    return tuple(-v for v in o)


# ---- linear constant-coefficient stencils: out[c] = sum_o w_o u[c+o];  adjoint: diffu[c] = sum_o w_o diffout[c-o] -------------
def _weights_c2(alpha=0.1):
    return {(0, 0): 1 - 4 * alpha, (1, 0): alpha, (-1, 0): alpha, (0, 1): alpha, (0, -1): alpha}


def _weights_c3(alpha=0.1):
    w = {(0, 0, 0): 1 - 6 * alpha}
    for ax in range(3):
        for s in (1, -1):
            o = [0, 0, 0]
            o[ax] = s
            w[tuple(o)] = alpha
    return w


def _weights_c4(weights=(0.4, 0.05, 0.02, 0.0075)):
    return {o: weights[sum(abs(v) for v in o)] for o in itertools.product((-1, 0, 1), repeat=3)}


def _linear(weights):
    gl = max(max(abs(v) for v in o) for o in weights)

    def forward(arrays, boundary='zeros', dtype=None):
        u = np.asarray(arrays['u'])
        acc = np.zeros(u.shape, dtype=np.float64)
        for o, w in weights.items():
            acc += w * shifted(u, o).astype(np.float64)
        return _finish({'out': acc}, gl, boundary, dtype or u.dtype)

    def backward(arrays, boundary='zeros', dtype=None):
        d = np.asarray(arrays['diffout'])
        acc = np.zeros(d.shape, dtype=np.float64)
        for o, w in weights.items():
            acc += w * shifted(d, _neg(o)).astype(np.float64)
        return _finish({'diffu': acc}, gl, boundary, dtype or d.dtype)
    return forward, backward


# ---- C1: z = x log(x y)  (README.rst:55-59; known adjoint README.rst:85-86) ---------------------------------------------------
def _c1_forward(arrays, boundary=None, dtype=None):
    x, y = (np.asarray(arrays[k]).astype(np.float64) for k in 'xy')
    return _finish({'z': x * np.log(x * y)}, 0, 'zeros', dtype or np.asarray(arrays['x']).dtype)


def _c1_backward(arrays, boundary=None, dtype=None):
    x, y, dz = (np.asarray(arrays[k]).astype(np.float64) for k in ('x', 'y', 'diffz'))
    return _finish({'diffx': dz * (np.log(x * y) + 1), 'diffy': dz * x / y}, 0, 'zeros',
                   dtype or np.asarray(arrays['x']).dtype)


# ---- C5: g = lam (u - f) - div( grad u / sqrt(|grad u|^2 + eps^2) ), forward differences for the gradient, backward
# differences for the divergence; images stacked along axis 0 (no coupling between images) ------------------------------------
_TV_READS = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 0, -1), (0, 1, -1), (0, -1, 1), (0, -1, 0)]


def _tv_rhs(U, f, lam, eps):
    """``U[o]``: the value of u at c + o for the seven offsets a cell reads."""
    def flux(dy, dx):
        ux = U[(0, dy, dx + 1)] - U[(0, dy, dx)]
        uy = U[(0, dy + 1, dx)] - U[(0, dy, dx)]
        n = np.sqrt(ux * ux + uy * uy + eps * eps)
        return ux / n, uy / n
    px_c, py_c = flux(0, 0)
    px_w, _ = flux(0, -1)
    _, py_s = flux(-1, 0)
    return lam * (U[(0, 0, 0)] - f) - ((px_c - px_w) + (py_c - py_s))


def _tv_forward(arrays, boundary='zeros', dtype=None, lam=1.0, eps=1e-3):
    u, f = np.asarray(arrays['u']), np.asarray(arrays['f'])
    U = {o: shifted(u, o).astype(np.float64) for o in _TV_READS}
    return _finish({'g': _tv_rhs(U, f.astype(np.float64), lam, eps)}, 1, boundary, dtype or u.dtype)


def _tv_coefficients(u, f, lam, eps):
    U = {o: shifted(u, o).astype(np.float64) for o in _TV_READS}
    f = np.asarray(f).astype(np.float64)
    h = 1e-30
    coef = {}
    for o in _TV_READS:
        Uc = {k: (v + 1j * h if k == o else v.astype(np.complex128)) for k, v in U.items()}
        coef[o] = np.imag(_tv_rhs(Uc, f, lam, eps)) / h
    return coef


def _tv_backward(arrays, boundary='zeros', dtype=None, lam=1.0, eps=1e-3, exact=False):
    """``exact=False``: the reference's rule (coefficient at the cell whose gradient is computed);  ``exact=True``: the
    transpose of the Jacobian (coefficient at the cell that performed the read) — 'zeros' mode only."""
    u, dg = np.asarray(arrays['u']), np.asarray(arrays['diffg'])
    coef = _tv_coefficients(u, arrays.get('f', np.zeros_like(u)), lam, eps)
    acc = np.zeros(u.shape, dtype=np.float64)
    for o in _TV_READS:
        if exact:
            acc += shifted(coef[o] * dg.astype(np.float64), _neg(o))
        else:
            acc += coef[o] * shifted(dg, _neg(o)).astype(np.float64)
    return _finish({'diffu': acc, 'difff': -lam * dg.astype(np.float64)}, 1, boundary, dtype or u.dtype)


_c2 = _linear(_weights_c2())
_c3 = _linear(_weights_c3())
_c4 = _linear(_weights_c4())

#: workload -> (forward, backward); both take ``(arrays, boundary, dtype)`` and return ``{output field name: array}``
HANDWRITTEN = {
    'c1': (_c1_forward, _c1_backward),
    'c2': _c2,
    'c3': _c3,
    'c4': _c4,
    'c5': (_tv_forward, _tv_backward),
}
