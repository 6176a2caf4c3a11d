"""Multi-GPU: slab-sharded forward + adjoint with NCCL halo exchange == unsharded, bit for bit (needs >= 2 GPUs)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.no_launch          # the launches happen in the rank processes this test starts
@pytest.mark.parametrize('name,bh', [('c3', 'zeros'), ('c3', 'none'), ('c4', 'zeros'), ('c2', 'zeros'), ('c5', 'zeros')])
def test_sharded_equals_unsharded(name, bh):
    n = _ngpu()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2 if n < 4 else 4
    port = 29700 + (hash((name, bh)) % 200)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                          '--master-addr', '127.0.0.1', '--master-port', str(port),
                          os.path.join(ROOT, 'scripts', 'check_slab.py'), name, bh],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    # the ranks print concurrently: two reports can share a line, so count verdicts, not lines
    assert out.stdout.count('[rank') >= world and out.stdout.count('IDENTICAL') == out.stdout.count('[rank')
    assert 'DIFFERENT' not in out.stdout


@pytest.mark.no_launch
@pytest.mark.parametrize('name,bh,steps', [('c3', 'zeros', 9), ('c3', 'none', 7), ('c4', 'zeros', 6)])
def test_peer_halos_equal_unsharded(name, bh, steps):
    """Peer halos (``SlabDataHandling(peer_halo=True)``: the stencil kernels stage their ghost planes by TMA from the
    neighbouring GPUs' arrays over NVLink, one launch per kernel, per-rank launch counters as the only ordering): time
    loops of single steps and fused pairs, and a sequence mixing peer launches with launches outside the protocol
    (``psad_peer_wait``, NCCL exchange), bit for bit against the unsharded kernels — many more launches than ranks, so a
    missed wait shows up as a difference."""
    n = _ngpu()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2 if n < 4 else 4
    port = 29400 + (hash((name, bh, 'peer')) % 200)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                          '--master-addr', '127.0.0.1', '--master-port', str(port),
                          os.path.join(ROOT, 'scripts', 'check_peer_halo.py'), name, bh, str(steps)],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count('[rank') >= 3 * world and out.stdout.count('IDENTICAL') == out.stdout.count('[rank')
    assert 'DIFFERENT' not in out.stdout


@pytest.mark.no_launch
@pytest.mark.parametrize('name,steps', [('c3', 5), ('c4', 4)])
def test_periodic_time_loops_wrap_around_between_gpus(name, steps):
    """Periodic along the decomposed axis (graph_datahandling.py:305-316): the first and the last rank exchange planes — with
    two ranks both neighbours are the same peer and the grouped ncclSend/ncclRecv must pair up in issue order.  Single
    steps and fused pairs, bit for bit against the whole field run as ONE periodic rank, that one against a torch.roll
    restatement (``scripts/check_periodic.py``)."""
    n = _ngpu()
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    world = 2 if n < 4 else 4
    port = 29250 + (hash((name, 'periodic')) % 100)
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                          '--master-addr', '127.0.0.1', '--master-port', str(port),
                          os.path.join(ROOT, 'scripts', 'check_periodic.py'), name, str(steps)],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count('[rank') == 2 * world and out.stdout.count('IDENTICAL') == 2 * world
    assert 'DIFFERENT' not in out.stdout
