"""Stand-ins for ``torch.cuda`` streams / events and pinned allocation.  TEST INFRASTRUCTURE: lets host code that is written
for CUDA tensors (three-stream end-to-end legs, host-streamed chunks) run on CPU tensors, with the kernels replayed by
``tests/replay_kernels.py``.  Everything executes synchronously in program order — one valid schedule of the stream
program — so this checks indexing, ranges, bookkeeping and plain Python errors, not the event dependencies themselves."""
import contextlib
import time

import torch


class FakeStream:
    cuda_stream = 0

    def __init__(self, device=None, **_):
        self.device = device

    def wait_event(self, event):
        pass

    def wait_stream(self, stream):
        pass

    def synchronize(self):
        pass

    def record_event(self, event=None):
        event = event or FakeEvent()
        event.record(self)
        return event


class FakeEvent:
    def __init__(self, enable_timing=False, **_):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def wait(self, stream=None):
        pass

    def synchronize(self):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class FakeGraph:
    """Stand-in for ``torch.cuda.CUDAGraph``: while ``capturing`` is set, ``tests/replay_kernels.ReplayKernel`` records its
    launches here instead of executing them (a capturing stream executes nothing either); ``replay()`` runs them on the
    tensors they were recorded with — pointers are baked in, exactly the property a real graph has."""
    capturing = None          # the graph being captured, if any
    replays = 0

    def __init__(self):
        self.ops = []

    def replay(self):
        type(self).replays += 1
        for fn in self.ops:
            fn()


@contextlib.contextmanager
def fake_graph_capture(graph, stream=None, **_):
    assert FakeGraph.capturing is None, 'nested capture'
    FakeGraph.capturing = graph
    try:
        yield
    finally:
        FakeGraph.capturing = None


@contextlib.contextmanager
def fake_cuda(graphs=False):
    """``graphs=True`` additionally makes ``torch.cuda.is_available()`` true, every tensor report ``is_cuda`` and
    ``torch.cuda.CUDAGraph`` / ``torch.cuda.graph`` the recording stand-ins above (for host code that only takes its CUDA-graph
    path on CUDA tensors)."""
    saved = {k: getattr(torch.cuda, k) for k in ('Stream', 'Event', 'current_stream', 'stream', 'synchronize', 'CUDAGraph',
                                                  'graph', 'is_available')}
    real_empty = torch.empty
    the_stream = FakeStream()

    @contextlib.contextmanager
    def stream_ctx(stream):
        yield

    def empty(*args, **kwargs):
        kwargs.pop('pin_memory', None)
        return real_empty(*args, **kwargs)

    torch.cuda.Stream, torch.cuda.Event = FakeStream, FakeEvent
    torch.cuda.current_stream = lambda device=None: the_stream
    torch.cuda.stream = stream_ctx
    torch.cuda.synchronize = lambda device=None: None
    torch.empty = empty
    if graphs:
        torch.cuda.CUDAGraph, torch.cuda.graph = FakeGraph, fake_graph_capture
        torch.cuda.is_available = lambda: True
        torch.Tensor.is_cuda = property(lambda self: True)      # shadows the C-level attribute while the context lasts
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(torch.cuda, k, v)
        torch.empty = real_empty
        if graphs:
            del torch.Tensor.is_cuda
