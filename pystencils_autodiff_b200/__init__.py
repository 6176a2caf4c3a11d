"""B200-native (sm_100a) execution backend for the ``torch_native`` path of pystencils_autodiff.

Public surface = the reference's (/root/reference/src/pystencils_autodiff/__init__.py:5-24) for this path, plus
the small stencil front end (``fields``, ``Assignment``, ``AssignmentCollection``, ``fd``) that stands in for the
third-party pystencils package.
"""
from . import backends, fd  # noqa: F401
from ._adjoint_field import AdjointField
from ._autodiff import (AutoDiffAstPair, AutoDiffBoundaryHandling, AutoDiffOp, DiffModes,
                        create_backward_assignments, get_jacobian_of_assignments)
from .assignment import Assignment, AssignmentCollection
from .field import Field, FieldType, fields
from .transformations import add_fixed_constant_boundary_handling

__version__ = '0.1.0'

__all__ = ['backends', 'fd', 'AdjointField', 'get_jacobian_of_assignments', 'create_backward_assignments',
           'AutoDiffOp', 'AutoDiffAstPair', 'DiffModes', 'AutoDiffBoundaryHandling', 'Assignment',
           'AssignmentCollection', 'Field', 'FieldType', 'fields', 'add_fixed_constant_boundary_handling',
           'show_code']


def show_code(op_or_kernel):
    """Specialised CUDA source of a kernel / op (reference: framework_integration/printer.py:200)."""
    from .emit import emit_kernel
    if hasattr(op_or_kernel, 'code'):
        return op_or_kernel.code
    return emit_kernel(op_or_kernel).source
