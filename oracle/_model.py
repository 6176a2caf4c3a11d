"""What the oracle needs to know about assignments — by DUCK TYPING, with no import from the product package.

TEST INFRASTRUCTURE.  The oracle must not share code with the thing it checks ("the oracle is not independent
of the product"): a mistake in the product's offset / index bookkeeping would otherwise be inherited.  Everything here reads
only the *published pystencils object protocol* the reference itself relies on:

* an **access** is a ``sympy.Symbol`` with ``.field``, ``.offsets`` (one integer per spatial axis, axis k <-> array dim k) and
  ``.index`` (``_autodiff.py:88-109`` reads exactly these);
* a **field** has ``.name``, ``.dtype.numpy_dtype``, ``.spatial_dimensions``, ``.index_shape`` and, when fixed-size,
  ``.spatial_shape`` (``backends/_torch_native.py:61-73,101-108``);
* an **assignment** has ``.lhs`` / ``.rhs``; a collection has ``.main_assignments`` / ``.subexpressions``
  (``_autodiff.py:35-38,75``);
* the ``'zeros'`` transform's node has ``.access``, ``.outofbounds_condition``, ``.outofbounds_value``
  (``transformations.py:26-30``).

So real pystencils objects, the product's front end and the stand-ins of the tests all evaluate through the same code.
"""
import numpy as np
import sympy as sp


def is_access(s):
    return isinstance(s, sp.Symbol) and hasattr(s, 'field') and hasattr(s, 'offsets') and hasattr(s, 'index')


def accesses_in(expr):
    return {s for s in expr.atoms(sp.Symbol) if is_access(s)}


def conditional_accesses_in(expr):
    return {c for c in expr.atoms(sp.Function) if hasattr(c, 'outofbounds_condition') and hasattr(c, 'access')}


def offsets_of(a):
    return tuple(int(o) for o in a.offsets)


def index_tail(a):
    return tuple(int(i) for i in a.index)


def ghost_width(a):
    """pystencils rule (Appendix A-3, from memory): the ghost width a kernel needs is the largest |offset| it uses."""
    return max([abs(o) for o in offsets_of(a)] + [0])


def field_dtype(f):
    dt = getattr(f.dtype, 'numpy_dtype', f.dtype)
    return np.dtype(dt)


class Collection:
    """Main assignments (lhs is a field access) and subexpressions, in order."""

    def __init__(self, main, subs):
        self.main_assignments = list(main)
        self.subexpressions = list(subs)

    @property
    def all_assignments(self):
        return self.subexpressions + self.main_assignments

    @property
    def free_symbols(self):
        defined = {a.lhs for a in self.all_assignments}
        out = set()
        for a in self.all_assignments:
            out |= a.rhs.free_symbols
        return out - defined

    def reads(self):
        """Every access on a right-hand side, including a ``+=`` form's read of its own output (_autodiff.py:110-113)."""
        out = set()
        for a in self.all_assignments:
            out |= accesses_in(a.rhs)
        return sorted(out, key=str)

    def writes(self):
        return [a.lhs for a in self.main_assignments]


def as_collection(assignments):
    if isinstance(assignments, Collection):
        return assignments
    if hasattr(assignments, 'main_assignments'):
        return Collection(assignments.main_assignments, getattr(assignments, 'subexpressions', []))
    if isinstance(assignments, dict):
        assignments = [sp.Eq(k, v, evaluate=False) for k, v in assignments.items()]
    main, subs = [], []
    for a in assignments:
        (main if is_access(a.lhs) else subs).append(a)
    return Collection(main, subs)


def spatial_shape(ac, arrays):
    acc = ac.reads() + ac.writes()
    for a in acc:
        if a.field.name in arrays:
            return tuple(np.asarray(arrays[a.field.name]).shape[:int(a.field.spatial_dimensions)])
    for a in acc:
        if getattr(a.field, 'has_fixed_shape', False):
            return tuple(int(s) for s in a.field.spatial_shape)
    raise ValueError('cannot infer the iteration shape')
