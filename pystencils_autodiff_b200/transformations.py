"""Symbolic ``'zeros'`` boundary handling.

Follows /root/reference/src/pystencils_autodiff/transformations.py:12-36: every relative field access ``a`` on a
right-hand side becomes ``ConditionalFieldAccess(a, out_of_bounds(ctr + offsets))`` whose value is 0 outside the
array; collections whose accesses all have zero offsets are returned unchanged (:18-19); the bounds come from
the spatial shape of one accessed field (:20).  The CUDA path does not evaluate this form — the kernels get the
same semantics from TMA out-of-bounds zero fill / an index predicate — but the lowering in ``ir.py`` recognises
it, and the oracle evaluates it literally, which is how the two are cross-checked.
"""
import sympy as sp

from .assignment import Assignment, AssignmentCollection, sympy_cse, coerce_assignments
from .field import Field, x_vector

__all__ = ['ConditionalFieldAccess', 'add_fixed_constant_boundary_handling']


class ConditionalFieldAccess(sp.Function):
    """``ConditionalFieldAccess(access, outofbounds_condition[, outofbounds_value])``"""
    nargs = (2, 3)

    @classmethod
    def eval(cls, *args):
        return None

    @property
    def access(self):
        return self.args[0]

    @property
    def outofbounds_condition(self):
        return self.args[1]

    @property
    def outofbounds_value(self):
        return self.args[2] if len(self.args) > 2 else sp.Integer(0)


def _out_of_bounds(offsets, counters, shape):
    """``ctr_k + o_k`` falls outside ``[0, N_k)`` on any axis."""
    per_axis = []
    for o, ctr, n in zip(offsets, counters, shape):
        pos = ctr + o
        per_axis.append(sp.Or(pos < 0, pos >= n))
    return sp.Or(*per_axis)


def add_fixed_constant_boundary_handling(assignments, with_cse=True):
    """Guard every relative read with "0 outside the array" (the reference's ``'zeros'`` boundary handling)."""
    collection = coerce_assignments(assignments)
    accesses = set()
    for a in collection.all_assignments:
        accesses |= a.atoms(Field.Access)
    if not any(o != 0 for acc in accesses for o in acc.offsets):
        return collection                      # pointwise kernels need no guard
    # all fields of a kernel share one spatial shape; take it from the first access in name order
    shape = min(accesses, key=str).field.spatial_shape
    counters = list(x_vector(len(shape)))

    def guard(expr):
        wrapped = {acc: ConditionalFieldAccess(acc, _out_of_bounds(acc.offsets, counters, shape))
                   for acc in expr.atoms(Field.Access) if not acc.is_absolute_access}
        return expr.subs(wrapped)

    guarded = [Assignment(a.lhs, guard(a.rhs)) for a in collection.all_assignments]
    result = AssignmentCollection([a for a in guarded if isinstance(a.lhs, Field.Access)],
                                  [a for a in guarded if not isinstance(a.lhs, Field.Access)])
    return sympy_cse(result) if with_cse else result
