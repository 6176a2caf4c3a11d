"""CPU fuzz: every variant of random stencils (forward and adjoint) is emitted and compiled by NVRTC for sm_100a."""
import pytest

import pystencils_autodiff_b200 as ps
from pystencils_autodiff_b200 import runtime
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from stencil_fuzz import random_stencil


@pytest.mark.parametrize('seed', range(8))
def test_random_stencils_compile(seed):
    asg, bh, shape, dtype = random_stencil(seed)
    op = ps.AutoDiffOp(asg, boundary_handling=bh, op_name='fuzz%d' % seed)
    for ir in (op.forward_ast_gpu, op.backward_ast_gpu):
        k = CompiledKernel(ir)
        assert 'generic' in k.variants
        for v in ('generic', 'march', 'march_nomask'):
            if v in k._emitted:
                ek = k.emitted(v)
                runtime.compile_source(ek.source, ek.cache_key, ek.options)
