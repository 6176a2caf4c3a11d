"""``Assignment`` / ``AssignmentCollection``: the container the reference differentiates.

Subset of pystencils' classes as used by the reference at
/root/reference/src/pystencils_autodiff/_autodiff.py:35-50 (``new_without_subexpressions``,
``main_assignments``, ``free_symbols``), :242-244, :271-275 (``free_fields`` / ``bound_fields``),
/root/reference/src/pystencils_autodiff/transformations.py:26-35 (``all_assignments``) and the printed
form shown in /root/reference/README.rst:62-68,81-86.
"""
from collections import OrderedDict

import sympy as sp

from .field import Field, coerce_access

__all__ = ['Assignment', 'AssignmentCollection', 'sympy_cse', 'sympy_cse_on_assignment_list', 'coerce_assignments']


class Assignment:
    """``lhs ← rhs``; lhs is a :class:`Field.Access` (main assignment) or a plain symbol (subexpression)."""

    def __init__(self, lhs, rhs):
        self.lhs = lhs
        self.rhs = sp.sympify(rhs)

    @property
    def free_symbols(self):
        return self.rhs.free_symbols

    def atoms(self, *types):
        res = set(self.rhs.atoms(*types))
        if isinstance(self.lhs, sp.Basic):
            res |= set(self.lhs.atoms(*types))
        return res

    def subs(self, *args, **kwargs):
        return Assignment(self.lhs, self.rhs.subs(*args, **kwargs))

    def __eq__(self, other):
        return isinstance(other, Assignment) and self.lhs == other.lhs and self.rhs == other.rhs

    def __hash__(self):
        return hash((self.lhs, self.rhs))

    def __str__(self):
        return '{lhs} ← {rhs}'.format(lhs=self.lhs, rhs=self.rhs)

    def __repr__(self):
        return 'Assignment(%s, %s)' % (sp.srepr(self.lhs) if not isinstance(self.lhs, Field.Access) else self.lhs,
                                       self.rhs)

    def __iter__(self):  # allows ``lhs, rhs = assignment``
        return iter((self.lhs, self.rhs))


def _to_assignment_list(assignments):
    if isinstance(assignments, dict):
        return [Assignment(k, v) for k, v in assignments.items()]
    if isinstance(assignments, AssignmentCollection):
        return list(assignments.all_assignments)
    out = []
    for a in assignments:
        if isinstance(a, Assignment):
            out.append(a)
        else:  # foreign assignment object with lhs/rhs
            out.append(Assignment(a.lhs, a.rhs))
    return out


class AssignmentCollection:
    def __init__(self, main_assignments, subexpressions=(), simplification_hints=None, subexpression_symbol_generator=None):
        if isinstance(main_assignments, dict):
            main_assignments = [Assignment(k, v) for k, v in main_assignments.items()]
        if isinstance(subexpressions, dict):
            subexpressions = [Assignment(k, v) for k, v in subexpressions.items()]
        self.main_assignments = _to_assignment_list(main_assignments)
        self.subexpressions = _to_assignment_list(subexpressions)
        self.simplification_hints = simplification_hints or {}

    # -- views ---------------------------------------------------------------------------------
    @property
    def all_assignments(self):
        return self.subexpressions + self.main_assignments

    def __iter__(self):
        return iter(self.all_assignments)

    def __len__(self):
        return len(self.all_assignments)

    @property
    def bound_symbols(self):
        return {a.lhs for a in self.all_assignments}

    @property
    def defined_symbols(self):
        return {a.lhs for a in self.main_assignments}

    @property
    def rhs_symbols(self):
        res = set()
        for a in self.all_assignments:
            res |= a.rhs.free_symbols
        return res

    @property
    def free_symbols(self):
        return self.rhs_symbols - self.bound_symbols

    @property
    def free_fields(self):
        """Fields that are read (pystencils semantics: fields of free Field.Access symbols)."""
        return {s.field for s in self.free_symbols if isinstance(s, Field.Access)}

    @property
    def bound_fields(self):
        return {s.field for s in self.bound_symbols if isinstance(s, Field.Access)}

    def atoms(self, *types):
        res = set()
        for a in self.all_assignments:
            res |= a.atoms(*types)
        return res

    # -- transformations -----------------------------------------------------------------------
    def new_without_subexpressions(self, subexpressions_to_keep=()):
        """Inline every subexpression into the main assignments."""
        if not self.subexpressions:
            return AssignmentCollection(list(self.main_assignments), [])
        subs_map = OrderedDict()
        for a in self.subexpressions:  # topological (definition) order
            subs_map[a.lhs] = a.rhs.xreplace(subs_map) if subs_map else a.rhs
        new_main = [Assignment(a.lhs, a.rhs.xreplace(subs_map)) for a in self.main_assignments]
        return AssignmentCollection(new_main, [])

    def copy(self, main_assignments=None, subexpressions=None):
        return AssignmentCollection(self.main_assignments if main_assignments is None else main_assignments,
                                    self.subexpressions if subexpressions is None else subexpressions)

    def subs(self, *args, **kwargs):
        return AssignmentCollection([a.subs(*args, **kwargs) for a in self.main_assignments],
                                    [a.subs(*args, **kwargs) for a in self.subexpressions])

    # -- identity / printing -------------------------------------------------------------------
    def __eq__(self, other):
        return isinstance(other, AssignmentCollection) and set(self.all_assignments) == set(other.all_assignments)

    def __hash__(self):
        return hash(frozenset(self.all_assignments))

    def __str__(self):
        result = 'Subexpressions:\n'
        for eq in self.subexpressions:
            result += '\t{eq}\n'.format(eq=eq)
        result += 'Main Assignments:\n'
        for eq in self.main_assignments:
            result += '\t{eq}\n'.format(eq=eq)
        return result

    def __repr__(self):
        return 'AssignmentCollection: ' + ', '.join(str(a.lhs) for a in self.main_assignments) + \
            ' <- f(' + ', '.join(sorted(str(s) for s in self.free_symbols)) + ')'


# ---------------------------------------------------------------------------------------------
def sympy_cse_on_assignment_list(assignments, symbol_prefix='xi'):
    """CSE over a list of assignments (reference call site: _autodiff.py:160, :415)."""
    assignments = _to_assignment_list(assignments)
    gen = (sp.Symbol('%s_%d' % (symbol_prefix, i)) for i in range(10 ** 6))
    replacements, reduced = sp.cse([a.rhs for a in assignments], symbols=gen, order='none')
    subexpr = [Assignment(s, e) for s, e in replacements]
    old_sub = [Assignment(a.lhs, r) for a, r in zip(assignments, reduced) if not isinstance(a.lhs, Field.Access)]
    main = [Assignment(a.lhs, r) for a, r in zip(assignments, reduced) if isinstance(a.lhs, Field.Access)]
    return _sort_topologically(subexpr + old_sub) + main


def sympy_cse(collection, symbol_prefix='xi'):
    """CSE over a collection (reference call site: transformations.py:33)."""
    res = sympy_cse_on_assignment_list(collection.all_assignments, symbol_prefix)
    return AssignmentCollection([a for a in res if isinstance(a.lhs, Field.Access)],
                                [a for a in res if not isinstance(a.lhs, Field.Access)])


def _sort_topologically(subexpressions):
    defined = {a.lhs for a in subexpressions}
    done, out = set(), []
    pending = list(subexpressions)
    while pending:
        progressed = False
        rest = []
        for a in pending:
            if (a.rhs.free_symbols & defined) <= done:
                out.append(a)
                done.add(a.lhs)
                progressed = True
            else:
                rest.append(a)
        if not progressed:
            raise ValueError('Cyclic subexpression definitions')
        pending = rest
    return out


def coerce_assignments(assignments):
    """Import a (possibly foreign, i.e. real-pystencils) list/collection of assignments.

    Every object exposing ``field/offsets/index`` is rebuilt as our :class:`Field.Access`; everything else
    stays a plain sympy expression.
    """
    if isinstance(assignments, AssignmentCollection):
        return assignments
    if hasattr(assignments, 'main_assignments') and hasattr(assignments, 'subexpressions'):
        main, sub = list(assignments.main_assignments), list(assignments.subexpressions)
    else:
        lst = _to_assignment_list(assignments)
        main = [a for a in lst if _is_access(a.lhs)]
        sub = [a for a in lst if not _is_access(a.lhs)]

    def conv_expr(e):
        e = sp.sympify(e)
        repl = {s: coerce_access(s) for s in e.free_symbols if _is_access(s) and not isinstance(s, Field.Access)}
        return e.xreplace(repl) if repl else e

    def conv(a):
        lhs = coerce_access(a.lhs) if _is_access(a.lhs) else a.lhs
        return Assignment(lhs, conv_expr(a.rhs))

    return AssignmentCollection([conv(a) for a in main], [conv(a) for a in sub])


def _is_access(s):
    return isinstance(s, Field.Access) or (hasattr(s, 'field') and hasattr(s, 'offsets') and hasattr(s, 'index'))
