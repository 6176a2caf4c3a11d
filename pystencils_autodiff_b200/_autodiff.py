"""Symbolic reverse-mode differentiation of stencil assignments and the operator factory.

Host-side mirror of /root/reference/src/pystencils_autodiff/_autodiff.py.  What is kept *semantically identical*
 because it defines what "the adjoint kernel" is:

* ``transposed-forward`` (TF-MAD, reference :22-173): gather-form adjoint
  ``diff_f[0] <- sum_{out, ra in reads(f)} d rhs_out / d ra * diff_out[-o_ra - o_lhs]`` — the coefficient is left
  un-shifted exactly like the reference (:106-109), ``+=`` form for time-constant fields (:110-117), contributions
  of several forward assignments summed (:119-123,:156), optional CSE (:158-163), exclusive-write check (:170).
* ``transposed`` (reference :354-437): scatter-form, one assignment per forward read access.
* field orderings: inputs/outputs sorted by name (:289-294).

What is different: there is no pystencils underneath.  ``forward_ast_gpu`` / ``backward_ast_gpu`` return a
:class:`~pystencils_autodiff_b200.ir.StencilKernelIR` (the thing the sm_100a emitter specialises) instead of a
pystencils AST, and there is deliberately **no CPU kernel**: ``forward_kernel_cpu`` raises.
"""
import collections
from enum import Enum
from typing import List

import sympy as sp

from ._adjoint_field import AdjointField
from .assignment import Assignment, AssignmentCollection, coerce_assignments, sympy_cse_on_assignment_list
from .backends import AVAILABLE_BACKENDS
from .field import Field
from .ir import lower_assignments
from .transformations import add_fixed_constant_boundary_handling

DEFAULT_OP_NAME = "autodiffop"

__all__ = ['AutoDiffOp', 'AutoDiffBoundaryHandling', 'DiffModes', 'create_backward_assignments',
           'AutoDiffAstPair', 'get_jacobian_of_assignments']


class AutoDiffBoundaryHandling(str, Enum):
    """In-kernel boundary strategies (reference :176-197).

    ======= =====================================================================
    NONE    interior cells only (margin = max |offset|), border cells are 0
    ZEROS   all cells; reads outside the array are 0 — forward *and* backward
    VALID   rejected, as in the reference (:246-247)
    ======= =====================================================================
    """
    NONE = None
    ZEROS = 'zeros'
    VALID = 'valid'


class DiffModes(str, Enum):
    TRANSPOSED = 'transposed'
    TF_MAD = 'transposed-forward'


def _is_constant(field, constant_fields):
    return field in constant_fields or field.name in constant_fields


def _inline_main(forward_assignments):
    """Main assignments with every subexpression inlined (reference :35-42, :365-372)."""
    fa = forward_assignments
    if hasattr(fa, 'new_without_subexpressions'):
        fa = fa.new_without_subexpressions()
    if hasattr(fa, 'main_assignments'):
        fa = fa.main_assignments
    return AssignmentCollection(list(fa), [])


def _finish(backward_list, do_cse):
    if do_cse:
        try:
            backward_list = sympy_cse_on_assignment_list(backward_list)
        except Exception:  # the reference swallows CSE failures (:162-165)
            pass
    main = [a for a in backward_list if isinstance(a.lhs, Field.Access)]
    sub = [a for a in backward_list if not isinstance(a.lhs, Field.Access)]
    return AssignmentCollection(main, sub)


def _shifted(expr, shift):
    """``expr`` evaluated at the cell ``shift`` away: every field access moves by ``shift``."""
    if not any(shift):
        return expr
    return expr.xreplace({a: a.get_shifted(*shift) for a in expr.atoms(Field.Access)})


def tf_mad_backward(forward_assignments, constant_fields=(), time_constant_fields=None, diff_fields_prefix='diff',
                    do_common_subexpression_elimination=True, adjoint_mode='reference'):
    """Gather-form adjoint.  Returns ``(backward_collection, field_map, read_accesses, write_accesses)``.

    ``adjoint_mode='reference'`` restates the reference's rule (:104-109): the contribution of a read ``f[o]`` in
    ``out[l] = rhs`` to ``diff_f[0]`` is ``(d rhs / d f[o]) * diff_out[-o - l]`` with the coefficient left AT THE CENTRE
    CELL — exact for coefficients that do not depend on fields (linear stencils) and for centre reads, not otherwise
    (SURVEY.md Appendix B-1).  ``'exact'`` is the transpose of the Jacobian: cell ``p`` of ``f`` is read by the evaluation
    at cell ``c = p - o``, which writes ``out[c + l]``, so the contribution is
    ``(d rhs / d f[o])(evaluated at p - o) * diff_out[-o + l]`` — the coefficient's own accesses shift by ``-o``.  It is
    what ``torch.autograd.gradcheck`` expects for non-linear stencils with ``boundary_handling='zeros'``; with interior
    iteration (``None``) it additionally needs upstream gradients that vanish on the un-iterated border."""
    if adjoint_mode not in ('reference', 'exact'):
        raise ValueError("adjoint_mode must be 'reference' or 'exact'")
    exact = adjoint_mode == 'exact'
    fwd = _inline_main(forward_assignments)
    reads = sorted([s for s in fwd.free_symbols if isinstance(s, Field.Access)], key=str)
    writes = [a.lhs for a in fwd.main_assignments]
    if not writes:
        raise AssertionError('No write accesses found')
    if not all(isinstance(w, Field.Access) for w in writes):
        raise AssertionError('Please check if your assignments are a AssignmentCollection or main_assignments only')

    in_fields = sorted({a.field for a in reads}, key=str)
    adj_in = {f: AdjointField(f, diff_fields_prefix) for f in in_fields if not _is_constant(f, constant_fields)}
    adj_out = {f: AdjointField(f, diff_fields_prefix) for f in {w.field for w in writes}}
    accumulate = set(time_constant_fields) if time_constant_fields is not None else set()

    contributions = collections.OrderedDict()
    for fa in fwd.main_assignments:
        lhs, rhs = fa.lhs, fa.rhs
        d_out = adj_out[lhs.field]
        for f in in_fields:
            if f not in adj_in:
                continue
            d_in = adj_in[f]
            if d_in.index_dimensions == 0:
                total = sp.Integer(0)
                for ra in reads:
                    if ra.field != f:
                        continue
                    coef = sp.diff(rhs, ra)
                    if exact:
                        coef = _shifted(coef, tuple(-o for o in ra.offsets))
                        flipped = tuple(-o + l for o, l in zip(ra.offsets, lhs.offsets))
                    else:
                        flipped = tuple(-o - l for o, l in zip(ra.offsets, lhs.offsets))
                    total += coef * d_out[flipped](*lhs.index)
                targets = {d_in.center(): total}
            elif d_in.index_dimensions == 1:
                # one adjoint component per index of the input field (the reference's loop :138-152 keeps only
                # the last component — SURVEY.md Appendix B-6; here every component is emitted)
                targets = collections.OrderedDict()
                for ra in reads:
                    if ra.field != f:
                        continue
                    coef = sp.diff(rhs, ra)
                    if exact:
                        coef = _shifted(coef, tuple(-o for o in ra.offsets))
                        flipped = tuple(-o + l for o, l in zip(ra.offsets, lhs.offsets))
                    else:
                        flipped = tuple(-o - l for o, l in zip(ra.offsets, lhs.offsets))
                    key = d_in.center.at_index(*ra.index)
                    targets[key] = targets.get(key, sp.Integer(0)) + coef * d_out[flipped](*lhs.index)
            else:
                raise NotImplementedError()
            for target, total in targets.items():
                if f in accumulate or f.name in accumulate:
                    total = target + total
                contributions.setdefault(target, []).append(total)

    backward = [Assignment(k, sp.Add(*v)) for k, v in contributions.items()]
    result = _finish(backward, do_common_subexpression_elimination)
    if not _has_exclusive_writes(result):
        raise AssertionError("Backward assignments don't have exclusive writes!")
    return result, {**adj_in, **adj_out}, reads, writes


def transposed_backward(forward_assignments, constant_fields=(), time_constant_fields=None,
                        diff_fields_prefix='diff', do_common_subexpression_elimination=True):
    """Scatter-form adjoint: forward reads become backward writes (reference :354-437)."""
    fwd = _inline_main(forward_assignments)
    reads = sorted([s for s in fwd.free_symbols if isinstance(s, Field.Access)], key=str)
    writes = [a.lhs for a in fwd.main_assignments]
    if not all(isinstance(w, Field.Access) for w in writes):
        raise AssertionError('Please assure that you only assign to fields in your main_assignments!')
    adj_in = {f: AdjointField(f, diff_fields_prefix) for f in {a.field for a in reads}}
    adj_out = {f: AdjointField(f, diff_fields_prefix) for f in {a.field for a in writes}}
    d_writes = sp.Matrix([adj_out[w.field][w.offsets](*w.index) for w in writes])
    rhs_vec = sp.Matrix([a.rhs for a in fwd.main_assignments])
    accumulate = set(time_constant_fields) if time_constant_fields is not None else set()

    backward = []
    for ra in reads:
        if _is_constant(ra.field, constant_fields):
            continue
        lhs = adj_in[ra.field][ra.offsets](*ra.index)
        rhs = (rhs_vec.diff(ra).transpose() * d_writes)[0, 0]
        if ra.field in accumulate or ra.field.name in accumulate:
            rhs = lhs + rhs
        backward.append(Assignment(lhs, rhs))
    result = _finish(backward, do_common_subexpression_elimination)
    if not _has_exclusive_writes(result):
        raise AssertionError("Backward assignments don't have exclusive writes."
                             " You should consider using 'transposed-forward' mode for resolving those conflicts")
    return result, {**adj_in, **adj_out}, reads, writes


class AutoDiffOp:
    """Forward + adjoint kernels of a stencil ``AssignmentCollection`` (reference :209-709, same constructor)."""

    def __init__(self,
                 forward_assignments: List[Assignment],
                 op_name: str = DEFAULT_OP_NAME,
                 boundary_handling: AutoDiffBoundaryHandling = None,
                 time_constant_fields: List[Field] = None,
                 constant_fields: List[Field] = (),
                 diff_fields_prefix='diff',
                 do_common_subexpression_elimination=True,
                 diff_mode=DiffModes.TF_MAD,
                 backward_assignments=None,
                 **kwargs):
        diff_mode = DiffModes(diff_mode)
        self._additional_symbols = []
        if 'target' in kwargs:
            assert kwargs['target'].lower() in ['cpu', 'gpu'], "AutoDiffOp always supports both cpu and gpu"
            del kwargs['target']
        kwargs.pop('no_chaching', None)  # stray kwarg the reference forwards (:725)
        # not a reference argument: 'reference' (default, parity) or 'exact' (see tf_mad_backward); keyword-only, so the
        # reference's positional signature is unchanged
        self._adjoint_mode = kwargs.pop('adjoint_mode', 'reference')
        if self._adjoint_mode != 'reference' and diff_mode != DiffModes.TF_MAD:
            raise NotImplementedError("adjoint_mode='exact' is implemented for diff_mode='transposed-forward'")

        forward_assignments = coerce_assignments(forward_assignments)
        if boundary_handling == AutoDiffBoundaryHandling.VALID:
            raise NotImplementedError('there seems to be still a bug with valid. -> Use "zeros"')

        self._forward_assignments = forward_assignments
        # a fresh list: the reference appends to (and thereby grows) its mutable default argument (:251-252)
        self._constant_fields = list(constant_fields) + ['indexVector']
        self._time_constant_fields = time_constant_fields
        self._kwargs = kwargs
        self.op_name = op_name
        self._do_common_subexpression_elimination = do_common_subexpression_elimination
        self._boundary_handling = boundary_handling
        self._diff_mode = diff_mode
        self._forward_read_accesses = None
        self._forward_write_accesses = None
        self._backward_field_map = None
        self._forward_ir = None
        self._backward_ir = None
        self._fused_ir = None

        if backward_assignments:
            self._backward_assignments = coerce_assignments(backward_assignments)
        elif diff_mode == DiffModes.TF_MAD:
            (self._backward_assignments, self._backward_field_map,
             self._forward_read_accesses, self._forward_write_accesses) = tf_mad_backward(
                forward_assignments, self._constant_fields, time_constant_fields, diff_fields_prefix,
                do_common_subexpression_elimination, self._adjoint_mode)
        elif diff_mode == DiffModes.TRANSPOSED:
            (self._backward_assignments, self._backward_field_map,
             self._forward_read_accesses, self._forward_write_accesses) = transposed_backward(
                forward_assignments, self._constant_fields, time_constant_fields, diff_fields_prefix,
                do_common_subexpression_elimination)
        else:
            raise NotImplementedError()

        by_name = lambda x: str(x)  # noqa: E731
        self._forward_input_fields = sorted(forward_assignments.free_fields, key=by_name)
        self._forward_output_fields = sorted(forward_assignments.bound_fields, key=by_name)
        self._backward_input_fields = sorted(self._backward_assignments.free_fields, key=by_name)
        self._backward_output_fields = sorted(self._backward_assignments.bound_fields, key=by_name)
        if diff_mode == DiffModes.TRANSPOSED and not backward_assignments:
            # the reference derives the lists from its field map in this mode (:430-437): adjoints of the forward
            # outputs are "the" backward inputs (forward fields read by the coefficients are not listed) — kept,
            # but in sorted instead of Python-set order
            fmap = self._backward_field_map
            self._backward_input_fields = [fmap[f] for f in self._forward_output_fields]
            self._backward_output_fields = [fmap[f] for f in self._forward_input_fields
                                            if not _is_constant(f, self._constant_fields)]

    # -- dunder --------------------------------------------------------------------------------------------
    def __hash__(self):
        return hash((str(self.forward_assignments), str(self.backward_assignments), str(self.constant_fields)))

    def __repr__(self):
        def indent(s):
            return s.replace('\n', '\n    ')
        return 'Forward:\n    %s\nBackward:\n    %s\n' % (indent(str(self.forward_assignments)),
                                                          indent(str(self.backward_assignments)))

    __str__ = __repr__

    def __getstate__(self):
        return {'forward_assignments': self.forward_assignments,
                'backward_assignments': self.backward_assignments,
                'kwargs': self._kwargs,
                'op_name': self.op_name,
                'boundary_handling': self._boundary_handling,
                'constant_fields': self._constant_fields,
                'time_constant_fields': self._time_constant_fields}

    def __setstate__(self, state):
        self.__init__(state['forward_assignments'], op_name=state.get('op_name', ''),
                      boundary_handling=state.get('boundary_handling', 'zeros'),
                      time_constant_fields=state.get('time_constant_fields'),
                      constant_fields=[c for c in (state.get('constant_fields') or []) if c != 'indexVector'],
                      backward_assignments=state['backward_assignments'], **state['kwargs'])

    # -- symbolic views ------------------------------------------------------------------------------------
    @property
    def forward_assignments(self):
        return self._forward_assignments

    @property
    def backward_assignments(self):
        return self._backward_assignments

    def jacobian(self):
        """Jacobian of the forward assignments with respect to the forward read accesses"""
        return get_jacobian_of_assignments(self._forward_assignments, self._forward_read_accesses)

    @property
    def forward_write_accesses(self):
        return self._forward_write_accesses

    @property
    def forward_read_accesses(self):
        return self._forward_read_accesses

    @property
    def backward_write_accesses(self):
        return [a.lhs for a in self.backward_assignments.main_assignments]

    @property
    def backward_read_accesses(self):
        return [a for a in self.backward_assignments.free_symbols if isinstance(a, Field.Access)]

    @property
    def backward_input_fields(self):
        return self._backward_input_fields

    @property
    def backward_output_fields(self):
        return self._backward_output_fields

    @property
    def backward_fields(self):
        return self._backward_output_fields + self._backward_input_fields

    @property
    def forward_fields(self):
        return self._forward_output_fields + self._forward_input_fields

    @property
    def forward_input_fields(self):
        return self._forward_input_fields

    @property
    def forward_output_fields(self):
        return self._forward_output_fields

    @property
    def constant_fields(self):
        return self._constant_fields

    @property
    def time_constant_fields(self):
        return self._time_constant_fields

    @property
    def boundary_handling(self):
        return self._boundary_handling

    # -- lowered kernels (what the emitter specialises) ---------------------------------------------------------
    def _lower(self, assignments, suffix):
        kw = {k: v for k, v in self._kwargs.items() if k in ('ghost_layers', 'data_type', 'fast_math')}
        return lower_assignments(assignments, self._boundary_handling, self.op_name + suffix, **kw)

    @property
    def forward_ast_gpu(self):
        """Lowered forward kernel (reference :495-508 builds a pystencils GPU AST here)."""
        if self._forward_ir is None:
            self._forward_ir = self._lower(self._forward_assignments, '_forward_gpu')
        return self._forward_ir

    @property
    def backward_ast_gpu(self):
        """Lowered adjoint kernel (reference :528-542)."""
        assert self._backward_assignments, 'No backward assignments!'
        if self._backward_ir is None:
            self._backward_ir = self._lower(self._backward_assignments, '_backward_gpu')
        return self._backward_ir

    @property
    def fused_assignments(self):
        """Forward and adjoint assignments as ONE collection (forward subexpressions inlined).

        Valid whenever the upstream gradients ``diff<out>`` are known together with the forward inputs — a linear loss,
        a prescribed adjoint source, or evaluating operator and adjoint operator side by side.  One kernel then reads
        every field once: for a nonlinear stencil the forward inputs that the adjoint's coefficients need are not
        fetched a second time (README example: 32 -> 24 bytes per cell)."""
        fwd = self._forward_assignments.new_without_subexpressions()
        bwd = self._backward_assignments
        written = {a.lhs.field.name for a in fwd.main_assignments}
        if written & {f.name for f in bwd.free_fields}:
            raise NotImplementedError('the adjoint reads a forward output: forward and adjoint cannot be fused')
        return AssignmentCollection(list(fwd.main_assignments) + list(bwd.main_assignments), list(bwd.subexpressions))

    @property
    def fused_ast_gpu(self):
        """Lowered forward+adjoint kernel (no counterpart in the reference, which always launches two kernels)."""
        if getattr(self, '_fused_ir', None) is None:
            fw, bw = self.forward_ast_gpu, self.backward_ast_gpu
            if fw.boundary == 'none' and fw.ghost_layers != bw.ghost_layers:
                raise NotImplementedError('forward and adjoint iterate over different interiors: use two kernels')
            self._fused_ir = self._lower(self.fused_assignments, '_fused_gpu')
        return self._fused_ir

    @property
    def fused_kernel_gpu(self):
        from .backends._torch_native import compile_kernel
        return compile_kernel(self.fused_ast_gpu)

    def _no_cpu(self, *_, **__):
        raise NotImplementedError(
            'pystencils_autodiff_b200 is a CUDA-only (sm_100a) backend: there is no CPU kernel and no CPU '
            'fallback. Use backend="torch_native", use_cuda=True.')

    forward_ast_cpu = property(_no_cpu)
    backward_ast_cpu = property(_no_cpu)
    forward_kernel_cpu = property(_no_cpu)
    backward_kernel_cpu = property(_no_cpu)

    @property
    def forward_kernel_gpu(self):
        from .backends._torch_native import compile_kernel
        return compile_kernel(self.forward_ast_gpu)

    @property
    def backward_kernel_gpu(self):
        from .backends._torch_native import compile_kernel
        return compile_kernel(self.backward_ast_gpu)

    def _create_kernel(self, assignments, suffix, args, kwargs):
        """Reference :592-598: ``ps.create_kernel(assignments, *args, **kwargs).compile()`` on the RAW assignments (no
        boundary transform; ``ghost_layers`` / ``data_type`` as given, the interior inferred from the offsets otherwise).
        pystencils' ``target`` defaults to 'cpu', which this backend does not have: pass ``target='gpu'``."""
        from .backends._torch_native import compile_kernel
        kwargs = dict(kwargs)
        target = kwargs.pop('target', args[0] if args else 'cpu')
        if str(getattr(target, 'name', target)).lower() != 'gpu':
            self._no_cpu()
        kw = {k: v for k, v in kwargs.items() if k in ('ghost_layers', 'data_type', 'fast_math')}
        return compile_kernel(lower_assignments(assignments, None, self.op_name + suffix, **kw))

    def create_forward_kernel(self, *args, **kwargs):
        return self._create_kernel(self._forward_assignments, '_forward_custom_gpu', args, kwargs)

    def create_backward_kernel(self, *args, **kwargs):
        return self._create_kernel(self._backward_assignments, '_backward_custom_gpu', args, kwargs)

    def get_forward_kernel(self, is_gpu):
        return self.forward_kernel_gpu if is_gpu else self.forward_kernel_cpu

    def get_backward_kernel(self, is_gpu):
        return self.backward_kernel_gpu if is_gpu else self.backward_kernel_cpu

    def symbolic_boundary_handled_assignments(self):
        """(forward, backward) in the reference's ``ConditionalFieldAccess`` form (transformations.py:12-36)."""
        if self._boundary_handling == AutoDiffBoundaryHandling.ZEROS:
            return (add_fixed_constant_boundary_handling(self._forward_assignments),
                    add_fixed_constant_boundary_handling(self._backward_assignments))
        return self._forward_assignments, self._backward_assignments

    # -- op factories ------------------------------------------------------------------------------------------
    def create_torch_op(self, *args, **kwargs):
        return self.create_tensorflow_op(*args, backend='torch_native', **kwargs)

    def create_unrolled_torch_op(self, steps, op_name=None, tuning=None, fuse=None):
        """``Function`` applying this one-field stencil ``steps`` times (pairs of steps fused into one launch); see
        ``backends/_torch_native.create_unrolled_function``.  Not part of the reference API: its users chain
        ``op.apply`` calls, which still works here."""
        from .backends._torch_native import create_unrolled_function
        return create_unrolled_function(self, steps, op_name=op_name, tuning=tuning, fuse=fuse)

    def create_slab_torch_op(self, data_handling, **kwargs):
        """``Function`` of this op on one rank's slab of slab-decomposed fields (halo exchange in forward and on the
        upstream gradient in backward); see ``datahandling.create_slab_autograd_function``.  Not part of the reference
        API, which has no distributed path."""
        from .datahandling import create_slab_autograd_function
        return create_slab_autograd_function(self, data_handling, **kwargs)

    def create_slab_unrolled_torch_op(self, data_handling, steps, fuse=None, **kwargs):
        """``steps`` unrolled steps of this one-field linear stencil on one rank's slab as one ``Function``; see
        ``datahandling.create_slab_unrolled_function``."""
        from .datahandling import create_slab_unrolled_function
        return create_slab_unrolled_function(self, data_handling, steps, fuse=fuse, **kwargs)

    def create_tensorflow_op(self, inputfield_tensor_dict={}, forward_loop=None, backward_loop=None,
                             use_cuda=True, backend='tensorflow'):
        """Same dispatch signature as the reference (:611-616); only ``backend='torch_native', use_cuda=True`` exists."""
        backend = backend.lower()
        assert backend in AVAILABLE_BACKENDS, "\"{}\" is not a valid backend. Available backends: {}".format(
            backend, AVAILABLE_BACKENDS)
        if backend != 'torch_native':
            raise NotImplementedError(
                'backend=%r is not provided by pystencils_autodiff_b200; it replaces the torch_native path only'
                % backend)
        if not use_cuda:
            self._no_cpu()
        from .backends._torch_native import create_autograd_function
        return create_autograd_function(self, use_cuda,
                                        op_name=self.op_name if self.op_name != DEFAULT_OP_NAME else None)


def create_backward_assignments(forward_assignments,
                                diff_fields_prefix="diff",
                                time_constant_fields=[],
                                constant_fields=[],
                                diff_mode=DiffModes.TF_MAD,
                                do_common_sub_expression_elimination=True,
                                adjoint_mode='reference'):
    """Reference signature: _autodiff.py:713-718 (``adjoint_mode`` is this package's, see ``tf_mad_backward``)."""
    auto_diff = AutoDiffOp(forward_assignments,
                           diff_fields_prefix=diff_fields_prefix,
                           time_constant_fields=time_constant_fields,
                           constant_fields=constant_fields,
                           diff_mode=diff_mode,
                           do_common_subexpression_elimination=do_common_sub_expression_elimination,
                           adjoint_mode=adjoint_mode)
    return auto_diff.backward_assignments


class AutoDiffAstPair:
    """A (forward, backward) pair of already lowered kernels (reference :732-756); CUDA only."""

    def __init__(self, forward_ast, backward_ast, compilation_target='gpu'):
        if compilation_target != 'gpu':
            raise NotImplementedError('CUDA-only backend')
        self.forward_ast = forward_ast
        self.backward_ast = backward_ast
        self._forward_kernel = None
        self._backward_kernel = None

    def forward(self, *args, **kwargs):
        if self._forward_kernel is None:
            from .backends._torch_native import compile_kernel
            self._forward_kernel = compile_kernel(self.forward_ast)
        return self._forward_kernel(*args, **kwargs)

    def backward(self, *args, **kwargs):
        if self._backward_kernel is None:
            from .backends._torch_native import compile_kernel
            self._backward_kernel = compile_kernel(self.backward_ast)
        return self._backward_kernel(*args, **kwargs)

    __call__ = forward


def _has_exclusive_writes(assignment_collection):
    """No two main assignments write the same (field, index) (reference :759-779)."""
    seen = set()
    for a in assignment_collection.main_assignments:
        if isinstance(a.lhs, Field.Access):
            key = (a.lhs.field, a.lhs.index)
            if key in seen:
                return False
            seen.add(key)
    return True


def get_jacobian_of_assignments(assignments, diff_variables):
    """Jacobian of the right-hand sides w.r.t. ``diff_variables`` (reference :782-798)."""
    if hasattr(assignments, 'main_assignments'):
        assignments = assignments.main_assignments
    rhs = sp.Matrix([e.rhs for e in assignments])
    return rhs.jacobian(diff_variables)
