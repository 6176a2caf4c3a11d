"""Evaluation plans for sets of linear combinations that share inputs (the per-plane sums of a stencil).

A thread of the march kernel evaluates, for every cell it owns, one or more weighted sums over the elements of the
plane that just arrived.  Neighbouring cells (and different sums of the same cell) overlap heavily: for the 27-point
stencil the vertical pair ``u[y-1,x] + u[y+1,x]`` is used by the face sum of cell x and by the corner sums of cells
x-1 and x+1, for both the z and the z+-1 weights.  ``plan_linear`` finds such shared partial sums:

1. every target is split by coefficient: ``target = sum_w  w * S_w`` with ``S_w`` a plain sum of elements;
2. greedy common-pair elimination over all the ``S_w`` of all targets: the pair of addends that co-occurs most often
   becomes a temporary, until no pair occurs twice;
3. identical sums are evaluated once.

The result is deterministic (ties are broken by name) and independent of anything but the symbolic inputs, so every
instance of a kernel rounds identically.
"""
from collections import Counter, OrderedDict

import sympy as sp

__all__ = ['plan_linear', 'LinearPlan']


class LinearPlan:
    def __init__(self):
        self.temps = []       # [(name, a, b)]: name = a + b   (a, b: element symbols or temp names)
        self.sums = []        # [(name, [addends])]
        self.targets = []     # [(target key, [(coefficient expr, sum name or single addend)])]

    def op_count(self):
        adds = len(self.temps) + sum(max(0, len(a) - 1) for _, a in self.sums)
        return adds + sum(len(t) + max(0, len(t) - 1) for _, t in self.targets)


def _linear_terms(expr, elements):
    """expr -> {coefficient: [element, ...]} or None if expr is not a linear combination of ``elements``."""
    groups = OrderedDict()
    for term in sp.Add.make_args(sp.expand(expr)):
        coeff, rest = term.as_independent(*elements, as_Add=False)
        if rest not in elements:
            return None
        if coeff.has(*elements):
            return None
        groups.setdefault(coeff, []).append(rest)
    return groups


def plan_linear(targets, elements):
    """``targets``: [(key, expr)], ``elements``: set of element symbols.  Returns a :class:`LinearPlan` or None when a
    target is not linear in the elements."""
    elements = set(elements)
    split = []
    for key, expr in targets:
        g = _linear_terms(expr, elements)
        if g is None:
            return None
        split.append((key, g))
    # working sets of addend names
    sets = []     # [[names]] aligned with refs
    refs = []     # (target index, coefficient)
    for ti, (key, g) in enumerate(split):
        for coeff, elems in g.items():
            names = sorted(str(e) for e in elems)
            if len(set(names)) != len(names):
                return None
            sets.append(names)
            refs.append((ti, coeff))
    plan = LinearPlan()
    n_tmp = 0
    while True:
        pairs = Counter()
        for s_ in sets:
            if len(s_) >= 2:
                for i in range(len(s_)):
                    for j in range(i + 1, len(s_)):
                        pairs[(s_[i], s_[j])] += 1
        if not pairs:
            break
        best = max(pairs.items(), key=lambda kv: (kv[1], tuple(reversed(kv[0]))))
        (a, b), cnt = best
        if cnt < 2:
            break
        name = 'lt%d' % n_tmp
        n_tmp += 1
        plan.temps.append((name, a, b))
        for s_ in sets:
            if a in s_ and b in s_:
                s_.remove(a)
                s_.remove(b)
                s_.append(name)
                s_.sort()
    sum_names = {}
    for s_ in sets:
        key = tuple(s_)
        if len(s_) > 1 and key not in sum_names:
            sum_names[key] = 'ls%d' % len(sum_names)
            plan.sums.append((sum_names[key], list(s_)))
    per_target = [[] for _ in split]
    for s_, (ti, coeff) in zip(sets, refs):
        per_target[ti].append((coeff, sum_names[tuple(s_)] if len(s_) > 1 else s_[0]))
    for (key, _), lst in zip(split, per_target):
        plan.targets.append((key, sorted(lst, key=lambda cw: sp.default_sort_key(cw[0]))))
    return plan
