"""CPU oracle for the forward/adjoint stencil path.  TEST INFRASTRUCTURE ONLY.

``oracle.evaluate`` — numpy restatement (checker);  ``oracle.cgen`` — C/OpenMP restatement of the pystencils
CPU path (timing baseline + large-size checker).  PARITY UNPINNED at the pystencils boundary: see the header of
``oracle/evaluate.py``.  Nothing under ``pystencils_autodiff_b200/`` imports this package.
"""
from .evaluate import evaluate, evaluate_literal, evaluate_loops, forward_backward  # noqa: F401
