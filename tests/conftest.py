import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)')
    config.addinivalue_line('markers', 'no_launch: a gpu test that checks host behaviour only (no kernel launch expected)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu or config.pluginmanager.hasplugin('dryrun_plugin'):
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



#: GPU tests that check argument errors / host behaviour only and legitimately launch nothing
_NO_LAUNCH_OK = ('error', 'front_door', 'raises', 'rejects', 'picklable', 'show_code')


@pytest.fixture(autouse=True)
def _gpu_tests_really_launch(request):
    """A ``gpu``-marked test that passes without a single launch through the C ABI (``psad_kernel_launch``) would be a test
    of something other than the CUDA path: fail it.  (The CPU dry run of GPU test bodies lives in tests/dryrun_plugin.py,
    which this conftest never loads and which refuses to load next to a real GPU.)"""
    if 'gpu' not in request.keywords or request.config.pluginmanager.hasplugin('dryrun_plugin'):
        yield
        return
    from pystencils_autodiff_b200 import runtime
    n0 = runtime.launch_count()
    yield
    name = request.node.name.lower()
    if runtime.launch_count() == n0 and not any(k in name for k in _NO_LAUNCH_OK) \
            and request.node.get_closest_marker('no_launch') is None and not hasattr(request.node, '_skipped_by_test'):
        rep = getattr(request.node, 'rep_call', None)
        if rep is not None and rep.passed:
            pytest.fail('%s passed without launching a kernel through psad_kernel_launch' % request.node.nodeid)


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_makereport(item, call):
    outcome = yield
    rep = outcome.get_result()
    if rep.when == 'call':
        item.rep_call = rep
