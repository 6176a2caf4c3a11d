"""Random stencil generator shared by the CPU (compile-only) and GPU (parity) fuzz tests."""
import random

import sympy as sp

import pystencils_autodiff_b200 as ps


def random_stencil(seed, shape=None):
    rnd = random.Random(seed)
    ndim = rnd.choice([2, 3, 3])
    dtype = rnd.choice(['float32', 'float64'])
    if shape is None:
        shape = {2: (rnd.choice([17, 40]), rnd.choice([132, 128, 36])),
                 3: (rnd.choice([5, 9]), rnd.choice([11, 24]), rnd.choice([132, 64, 36]))}[ndim]
    n_in, n_out = rnd.choice([1, 2, 3]), rnd.choice([1, 1, 2])
    names = ['a', 'b', 'c'][:n_in] + ['p', 'q'][:n_out]
    flds = ps.fields('%s: %s[%s]' % (', '.join(names), dtype, ','.join(map(str, shape))))
    ins, outs = flds[:n_in], flds[n_in:]
    lim = {2: [(-2, 2), (-4, 4)], 3: [(-2, 2), (-2, 2), (-4, 4)]}[ndim]
    narrow = rnd.random() < 0.5

    def access():
        f = rnd.choice(ins)
        off = tuple(rnd.randint(lo, hi) if not narrow else rnd.randint(max(lo, -1), min(hi, 1)) for lo, hi in lim)
        return f[off]

    def term():
        kind = rnd.choice(['lin', 'lin', 'lin', 'prod', 'exp', 'sqrt', 'sin'])
        c = sp.Float(round(rnd.uniform(-1.5, 1.5), 3))
        if kind == 'lin':
            return c * access()
        if kind == 'prod':
            return c * access() * access()
        if kind == 'exp':
            return c * sp.exp(-access() ** 2)
        if kind == 'sqrt':
            return c * access() / sp.sqrt(access() ** 2 + 1)
        return c * sp.sin(access())

    asg = {}
    for o in outs:
        asg[o.center] = sp.Add(*[term() for _ in range(rnd.randint(2, 6))])
    bh = rnd.choice([None, 'zeros'])
    return ps.AssignmentCollection(asg), bh, shape, dtype
