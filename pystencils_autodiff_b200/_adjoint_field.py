"""``AdjointField``: the ``diff<name>`` companion of a forward field.

Mirrors /root/reference/src/pystencils_autodiff/_adjoint_field.py:9-30: name = prefix + forward name, same dtype /
layout / shape / strides, symbolic shape and stride symbols re-keyed to the adjoint's own name, LaTeX name
``\\hat{<forward>}`` (which is what the README prints as the lhs of backward assignments, README.rst:85-86).
"""
import sympy as sp

from .field import Field, FieldType

ADJOINT_FIELD_LATEX_HIGHLIGHT = r"\hat{%s}"


class AdjointField(Field):
    """Field representing adjoint variables to a Field representing the forward variables"""

    def __init__(self, forward_field, name_prefix='diff'):
        new_name = name_prefix + forward_field.name
        field_type = FieldType.GENERIC if forward_field.field_type != FieldType.BUFFER else FieldType.BUFFER
        super().__init__(new_name, field_type, forward_field.dtype, forward_field.layout,
                         forward_field.shape, forward_field.strides)
        self._index_dimensions = forward_field.index_dimensions
        self.corresponding_forward_field = forward_field
        self.name_prefix = name_prefix

        def rekey(sym, kind, i):
            if isinstance(sym, sp.Symbol):
                return sp.Symbol('_%s_%s_%d' % (kind, new_name, i), integer=True)
            return sym
        self.shape = tuple(rekey(s, 'size', i) for i, s in enumerate(self.shape))
        self.strides = tuple(rekey(s, 'stride', i) for i, s in enumerate(self.strides))

        if forward_field.latex_name:
            self.latex_name = ADJOINT_FIELD_LATEX_HIGHLIGHT % forward_field.latex_name
        else:
            self.latex_name = ADJOINT_FIELD_LATEX_HIGHLIGHT % forward_field.name
