import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu or os.environ.get('PSAD_REPLAY_GPU_TESTS'):
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session', autouse=True)
def _native_runtime_built():
    """The C-ABI library must exist for every test session; build it (g++ only) if a fresh checkout lacks it."""
    from pystencils_autodiff_b200 import runtime
    runtime.build_library()
    yield


@pytest.fixture(scope='session', autouse=True)
def _replay_gpu_tests_on_the_cpu():
    """``PSAD_REPLAY_GPU_TESTS=1 pytest -m gpu tests/test_gpu_zz_*.py``: a DRY RUN of GPU test bodies on a machine without
    a GPU — every kernel launch becomes a CPU replay of the emitted kernel (tests/replay_kernels.py), ``Tensor.cuda()`` /
    ``.to('cuda')`` the identity, streams and events stand-ins (tests/fake_cuda.py).  It finds Python-level mistakes in
    tests written without a GPU at hand; it proves nothing about the GPU and is never used by the driver's test runs."""
    if not os.environ.get('PSAD_REPLAY_GPU_TESTS'):
        yield
        return
    import contextlib
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import fake_cuda
    import replay_kernels
    from pystencils_autodiff_b200.backends import _torch_native
    from pystencils_autodiff_b200 import datahandling
    real = dict(call=_torch_native.CompiledKernel.__call__, cuda=torch.Tensor.cuda, to=torch.Tensor.to,
                count=torch.cuda.device_count)

    def is_cuda_dev(a):
        return (isinstance(a, str) and a.startswith('cuda')) or (isinstance(a, torch.device) and a.type == 'cuda')

    def to(self, *args, **kwargs):
        args = tuple('cpu' if is_cuda_dev(a) else a for a in args)
        if is_cuda_dev(kwargs.get('device')):
            kwargs['device'] = 'cpu'
        return real['to'](self, *args, **kwargs)

    factories = {}
    for fname in ('empty', 'zeros', 'ones', 'full', 'rand', 'randn', 'tensor', 'arange', 'empty_like', 'zeros_like',
                  'full_like', 'ones_like', 'rand_like', 'randn_like'):
        factories[fname] = getattr(torch, fname)

        def make(fn):
            def wrapped(*args, **kwargs):
                if is_cuda_dev(kwargs.get('device')):
                    kwargs['device'] = 'cpu'
                kwargs.pop('pin_memory', None)
                return fn(*args, **kwargs)
            return wrapped
        setattr(torch, fname, make(factories[fname]))
    real['pin'] = torch.Tensor.pin_memory
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    real['gen'] = torch.Generator

    class CpuGenerator(torch.Generator):
        def __new__(cls, device=None):
            return real['gen'].__new__(cls, 'cpu')
    torch.Generator = CpuGenerator

    _torch_native.CompiledKernel.__call__ = replay_kernels.ReplayKernel.__call__     # every instance, isinstance intact
    _torch_native.CompiledKernel.launches = []
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.to = to
    torch.cuda.device_count = lambda: 1
    torch.Tensor.is_cuda = property(lambda self: True)      # shadows the C-level attribute for the dry run only
    real_init = datahandling.SlabDataHandling.__init__

    def init_on_cpu(self, domain_size, rank=0, world_size=1, default_ghost_layers=1, device=None, backend='nccl', group=None):
        real_init(self, domain_size, rank, world_size, default_ghost_layers, 'cpu', 'torch', group)
    datahandling.SlabDataHandling.__init__ = init_on_cpu
    with fake_cuda.fake_cuda():
        try:
            yield
        finally:
            _torch_native.CompiledKernel.__call__ = real['call']
            torch.Tensor.cuda, torch.Tensor.to = real['cuda'], real['to']
            torch.cuda.device_count = real['count']
            del torch.Tensor.is_cuda
            for fname, fn in factories.items():
                setattr(torch, fname, fn)
            torch.Tensor.pin_memory = real['pin']
            torch.Generator = real['gen']
            datahandling.SlabDataHandling.__init__ = real_init
