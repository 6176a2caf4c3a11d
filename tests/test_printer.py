"""The CUDA expression printer, checked on the CPU: printed expressions are compiled with g++ (tiny prelude for the
device helpers) and compared with sympy's numeric evaluation."""
import random
import subprocess

import numpy as np
import pytest
import sympy as sp

from pystencils_autodiff_b200.emit import _CudaPrinter

PRELUDE = r'''
#include <cmath>
#include <cstdio>
template <int N, typename T> static T psad_ipow(T x) { T r = x; for (int i = 1; i < N; ++i) r *= x; return r; }
static float psad_rsqrt(float x) { return 1.0f / std::sqrt(x); }
static double psad_rsqrt(double x) { return 1.0 / std::sqrt(x); }
'''

x, y, z, w = sp.symbols('x y z w')
EXPRS = [
    0.1 * x + 0.4 * y + 0.1 * z - w,
    x * sp.log(x * y),
    (x - y) / sp.sqrt((x - y) ** 2 + (z - y) ** 2 + 1e-6),
    -(x - y) / sp.sqrt(x ** 2 + 1) + z * (x * y + 1) ** sp.Rational(-3, 2),
    x ** sp.Rational(3, 2) + y ** sp.Rational(5, 2) - 1 / z,
    sp.exp(-x ** 2) * sp.sin(y) + sp.cos(z) / (1 + w ** 2),
    x / y / z + x ** -2 - 3 * y ** 3 + sp.Rational(1, 3) * z,
    sp.Piecewise((x, x > y), (y * z, True)) + sp.Abs(w - 2) + sp.Max(x, z) - sp.Min(y, w),
    sum(sp.Float(0.01 * (i + 1)) * s for i, s in enumerate([x, y, z, w, x * y, y * z, z * w, x * w, x * z, y * w, x ** 2, y ** 2])),
    2 * x * y * z * w - (x + y) * (z - w) / (x * y + 2),
]


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_printed_expressions_evaluate_like_sympy(dtype, tmp_path):
    T = 'float' if dtype == np.float32 else 'double'
    pr = _CudaPrinter(dtype)
    random.seed(1)
    vals = {s: random.uniform(0.5, 2.0) for s in (x, y, z, w)}
    body = []
    for i, e in enumerate(EXPRS):
        body.append('  { const %s r = %s; std::printf("%%.17g\\n", (double)r); }' % (T, pr.doprint(e)))
    src = PRELUDE + 'int main() {\n' + ''.join('  const %s %s = %.17g;\n' % (T, s, v) for s, v in vals.items()) \
        + '\n'.join(body) + '\n  return 0;\n}\n'
    cpp = tmp_path / 'p.cpp'
    cpp.write_text(src)
    exe = tmp_path / 'p'
    subprocess.check_call(['g++', '-O1', '-ffp-contract=off', '-o', str(exe), str(cpp)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    tol = 2e-6 if dtype == np.float32 else 1e-13
    for e, o in zip(EXPRS, out):
        ref = float(e.subs(vals).evalf(30))
        assert abs(float(o) - ref) <= tol * max(1.0, abs(ref)), (e, o, ref)


def test_no_pow_calls_or_divisions_by_square_roots():
    pr = _CudaPrinter(np.float32)
    s = pr.doprint(-(x - y) / sp.sqrt((x - y) ** 2 + 1e-6) + z * w ** sp.Rational(-3, 2))
    assert 'psad_rsqrt(' in s and 'powf' not in s and '/sqrtf' not in s and 'psad_ipow<3>' in s
    assert pr.doprint(x ** sp.Rational(3, 2)) == '(sqrtf(x)*x)'
    assert pr.doprint(x / y) == 'x/y'
