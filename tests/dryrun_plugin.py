"""CPU DRY RUN of GPU test bodies — an OPT-IN pytest plugin, never loaded by ``pytest -m gpu``:

    python -m pytest -p dryrun_plugin -m gpu tests/test_gpu_zz_golden.py        (with tests/ on PYTHONPATH)

Every kernel launch becomes a CPU replay of the emitted kernel (tests/replay_kernels.py), ``Tensor.cuda()`` / ``.to('cuda')``
the identity, streams and events stand-ins (tests/fake_cuda.py).  It finds Python-level mistakes in tests written without a
GPU at hand and re-checks host logic; it proves nothing about the GPU.  It refuses to load when a CUDA device is present,
so a GPU run can never silently become a CPU run."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    import torch
    if torch.cuda.is_available():
        raise pytest.UsageError('dryrun_plugin replaces kernel launches by CPU replays: not allowed on a machine with a GPU')


@pytest.fixture(scope='session', autouse=True)
def _replay_gpu_tests_on_the_cpu():
    """See the module docstring."""
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

    def init_on_cpu(self, domain_size, rank=0, world_size=1, default_ghost_layers=1, device=None, backend='nccl', group=None,
                    periodic=False, peer_halo=False):
        real_init(self, domain_size, rank, world_size, default_ghost_layers, 'cpu', 'torch', group, periodic=periodic)
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
