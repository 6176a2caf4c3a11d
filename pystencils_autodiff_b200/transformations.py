"""Symbolic ``'zeros'`` boundary handling.

Follows /root/reference/src/pystencils_autodiff/transformations.py:12-36: every relative field access ``a`` on a
right-hand side becomes ``ConditionalFieldAccess(a, out_of_bounds(ctr + offsets))`` whose value is 0 outside the
array; collections whose accesses all have zero offsets are returned unchanged (:18-19); the bounds come from
the spatial shape of one accessed field (:20).  The CUDA path does not evaluate this form — the kernels get the
same semantics from TMA out-of-bounds zero fill / an index predicate — but the lowering in ``ir.py`` recognises
it, and the oracle evaluates it literally, which is how the two are cross-checked.
"""
import itertools

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


def add_fixed_constant_boundary_handling(assignments, with_cse=True):
    assignments = coerce_assignments(assignments)
    field_accesses = set().union(itertools.chain.from_iterable(
        [a.atoms(Field.Access) for a in assignments]))

    if all(all(o == 0 for o in a.offsets) for a in field_accesses):
        return assignments
    common_shape = sorted(field_accesses, key=str)[0].field.spatial_shape
    ndim = len(common_shape)

    def is_out_of_bound(access, shape):
        return sp.Or(*[sp.Or(a < 0, a >= s) for a, s in zip(access, shape)])

    safe_assignments = [Assignment(
        assignment.lhs, assignment.rhs.subs({
            a: ConditionalFieldAccess(a, is_out_of_bound(sp.Matrix(a.offsets) + x_vector(ndim), common_shape))
            for a in assignment.rhs.atoms(Field.Access) if not a.is_absolute_access
        })) for assignment in assignments.all_assignments]

    main = [a for a in safe_assignments if isinstance(a.lhs, Field.Access)]
    sub = [a for a in safe_assignments if not isinstance(a.lhs, Field.Access)]
    result = AssignmentCollection(main, sub)
    if with_cse:
        result = sympy_cse(result)
    return result
