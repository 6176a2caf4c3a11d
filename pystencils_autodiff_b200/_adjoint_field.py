"""``AdjointField``: the ``diff<name>`` companion of a forward field.

Mirrors /root/reference/src/pystencils_autodiff/_adjoint_field.py:9-30: name = prefix + forward name, same dtype /
layout / shape / strides, symbolic shape and stride symbols re-keyed to the adjoint's own name, LaTeX name
``\\hat{<forward>}`` (which is what the README prints as the lhs of backward assignments, README.rst:85-86).
"""
import sympy as sp

from .field import Field, FieldType

ADJOINT_FIELD_LATEX_HIGHLIGHT = r"\hat{%s}"


class AdjointField(Field):
    """The adjoint companion ``<prefix><name>`` of a forward field: same element type, layout, shape and strides."""

    def __init__(self, forward_field, name_prefix='diff'):
        adjoint_name = name_prefix + forward_field.name
        kind = FieldType.BUFFER if forward_field.field_type == FieldType.BUFFER else FieldType.GENERIC
        super().__init__(adjoint_name, kind, forward_field.dtype, forward_field.layout, forward_field.shape,
                         forward_field.strides)
        self._index_dimensions = forward_field.index_dimensions
        self.corresponding_forward_field = forward_field
        self.name_prefix = name_prefix

        # symbolic extents / strides must not refer to the forward field, which a backward kernel may not receive
        def own(sym, what, axis):
            return sp.Symbol('_%s_%s_%d' % (what, adjoint_name, axis), integer=True) if isinstance(sym, sp.Symbol) else sym

        self.shape = tuple(own(s, 'size', i) for i, s in enumerate(self.shape))
        self.strides = tuple(own(s, 'stride', i) for i, s in enumerate(self.strides))
        self.latex_name = ADJOINT_FIELD_LATEX_HIGHLIGHT % (forward_field.latex_name or forward_field.name)
