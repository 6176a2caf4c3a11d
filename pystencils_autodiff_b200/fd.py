"""Finite-difference helpers ``Diff`` and ``Discretization2ndOrder``.

The reference's tests build their stencils with ``ps.fd.Diff(f, k)`` and
``ps.fd.Discretization2ndOrder(dx=1)`` (/root/reference/tests/test_tfmad.py:16-19,195-200).  Only what those call
sites need is provided: first derivatives of field accesses, discretised with the second-order central
difference ``(f[+e_k] - f[-e_k]) / (2 dx)``; nested ``Diff`` (second derivatives) are discretised by applying
the rule recursively.
"""
import sympy as sp

from .field import Field

__all__ = ['Diff', 'Discretization2ndOrder']


class Diff(sp.Function):
    """``Diff(arg, target)``: derivative of ``arg`` along spatial coordinate ``target``."""
    nargs = (2,)
    is_commutative = True

    @classmethod
    def eval(cls, arg, target):
        return None

    def __new__(cls, arg, target=-1, **kwargs):
        if isinstance(arg, Field):
            arg = arg.center
        return sp.Function.__new__(cls, sp.sympify(arg), sp.Integer(target), **kwargs)

    @property
    def arg(self):
        return self.args[0]

    @property
    def target(self):
        return int(self.args[1])


def _shift(expr, axis, amount):
    repl = {}
    for a in expr.atoms(Field.Access):
        repl[a] = a.neighbor(axis, amount)
    return expr.xreplace(repl)


class Discretization2ndOrder:
    def __init__(self, dx=sp.Symbol('dx'), dt=sp.Symbol('dt')):
        self.dx = dx
        self.dt = dt

    def _discretize_diff(self, d):
        inner = self(d.arg)
        k = d.target
        return (_shift(inner, k, 1) - _shift(inner, k, -1)) / (2 * self.dx)

    def __call__(self, expr):
        expr = sp.sympify(expr)
        if not expr.atoms(Diff):
            return expr

        def rec(e):  # outermost first: replace each top-level Diff by its discretisation
            if isinstance(e, Diff):
                return self._discretize_diff(e)
            if not e.args:
                return e
            return e.func(*[rec(a) for a in e.args])
        return rec(expr)
