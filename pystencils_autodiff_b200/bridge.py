"""Adapter for wiring this backend into the reference's dispatcher.

``from_reference_op`` takes the reference's ``AutoDiffOp`` (pystencils objects inside) and rebuilds it on this
package's types, keeping the reference's *own* backward assignments so that the adjoint that runs is exactly the
one the reference derived (/root/reference/src/pystencils_autodiff/_autodiff.py:280-294).
"""
from ._autodiff import AutoDiffOp


def from_reference_op(ref_op):
    constant = [f for f in (ref_op.constant_fields or []) if f != 'indexVector']
    return AutoDiffOp(ref_op.forward_assignments,
                      op_name=ref_op.op_name,
                      boundary_handling=getattr(ref_op, '_boundary_handling', None),
                      constant_fields=constant,
                      time_constant_fields=ref_op.time_constant_fields,
                      backward_assignments=ref_op.backward_assignments)
