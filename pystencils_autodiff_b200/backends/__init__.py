"""Backend names accepted by ``AutoDiffOp.create_tensorflow_op`` (reference: backends/__init__.py:9).

Only ``'torch_native'`` has an implementation in this package; the other names are accepted by the dispatcher
for signature compatibility and raise ``NotImplementedError``.
"""

AVAILABLE_BACKENDS = ['tensorflow', 'torch', 'tensorflow_native', 'torch_native']
