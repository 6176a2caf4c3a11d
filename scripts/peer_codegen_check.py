"""One GPU: is the peer-halo instance of a march kernel (PSAD_PEER) slower than the plain instance when it has NO neighbours?
Separates the cost of the generated code (three sets of tensor maps, source selection in the producer) from
the cost of actually talking to a neighbour.  Also a self-neighbour mode: the 'neighbour' arrays are this GPU's own arrays
(ghost planes are then read through the second / third tensor map from local memory, flags are local words).

    python scripts/peer_codegen_check.py [c3|c4] [planes]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pystencils_autodiff_b200 import runtime
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel
from pystencils_autodiff_b200.configs import CONFIG_SHAPES, make_config
from pystencils_autodiff_b200.datahandling import slab_ranges


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    name = args[0] if args else 'c3'
    shape = list(CONFIG_SHAPES[name]['shape'])
    if len(args) > 1:
        shape[0] = int(args[1])
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    g = 1
    local = (shape[0] + 2 * g,) + tuple(shape[1:])
    op = make_config(name, shape=local, boundary_handling='zeros')
    out = {'workload': name, 'owned_shape': shape}
    out['nvrtc_extra'] = os.environ.get('PSAD_NVRTC_EXTRA', '')
    for which, ir in (('forward', op.forward_ast_gpu), ('adjoint', op.backward_ast_gpu)):
        if '--quick' in sys.argv and which == 'adjoint':
            continue
        k = CompiledKernel(ir)
        dt = getattr(torch, str(ir.input_fields[0].dtype.numpy_dtype))
        arrays = {f.name: torch.randn(local, device=dev, dtype=dt) for f in k.fields}
        whole = slab_ranges((shape[0],) + tuple(shape[1:]), 0, shape[0], g, False, False, 'zeros', ir.ghost_layers, 3)[0]
        flags = torch.zeros(8, dtype=torch.int32, device=dev)
        flags[0] = 1 << 30      # "the neighbour" is far ahead: no waiting
        tensors = [arrays[f.name] for f in k.fields]

        def make_peer(neighbours):
            p = runtime.Peer()
            if neighbours:
                for i, t in enumerate(tensors):
                    p.lo_ptr[i] = p.hi_ptr[i] = t.data_ptr()
                p.lo_planes = p.hi_planes = local[0]
                p.flag_lo = p.flag_hi = flags.data_ptr()
            p.error_flag = flags.data_ptr() + 4
            p.ghost_planes = g
            p.expect = 0
            return p

        modes = {'plain': None, 'peer_no_neighbours': make_peer(False), 'peer_self_neighbour': make_peer(True)}
        res = {}
        for rep in range(2):
            for mode, peer in modes.items():
                kw = dict(arrays, _range=whole)
                if peer is not None:
                    kw['_peer'] = peer
                for _ in range(3):
                    k(**kw)
                torch.cuda.synchronize()
                time.sleep(1.5)          # every timed region starts from an idle GPU (power cap)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(20):
                    k(**kw)
                b.record()
                torch.cuda.synchronize()
                res.setdefault(mode, []).append(round(a.elapsed_time(b) / 20, 4))
                res[mode + '_instance'] = k.last_instance
        out[which] = res
        del arrays, tensors
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
