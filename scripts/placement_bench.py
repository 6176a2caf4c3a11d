"""Why is a kernel slower inside bench.py's step than stand-alone?  Separates three candidates on one GPU:

  (a) duration / power: the same launch repeated 20 vs 400 times, per-launch times from events on every launch;
  (b) pairing: forward u->out alone, ping-pong u<->out, and the bench's forward + adjoint over four arrays;
  (c) placement: input and output carved out of ONE allocation at a sweep of relative byte offsets.

    python scripts/placement_bench.py c4 [c3]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel, numpy_dtype_to_torch
from pystencils_autodiff_b200.configs import CONFIG_SHAPES, make_config


def per_launch(fn, iters, warm=5):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    evs[0].record()
    for i in range(iters):
        fn(i)
        evs[i + 1].record()
    torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(iters)]
    return ts


def stats(ts):
    s = sorted(ts)
    n = len(s)
    return dict(median=s[n // 2], best=s[0], p90=s[int(n * 0.9)], first10=sum(ts[:10]) / min(10, n), last10=sum(ts[-10:]) / min(10, n),
                total_over_n=sum(ts) / n)


def main():
    dev = torch.device('cuda:0')
    out = {}
    for name in sys.argv[1:] or ['c4']:
        shape = CONFIG_SHAPES[name]['shape']
        op = make_config(name, shape=shape, boundary_handling='zeros')
        fk, bk = CompiledKernel(op.forward_ast_gpu), CompiledKernel(op.backward_ast_gpu)
        dt = numpy_dtype_to_torch(fk.fields[0].dtype.numpy_dtype)
        fin, fout = op.forward_ast_gpu.input_fields[0].name, op.forward_ast_gpu.output_fields[0].name
        bin_, bout = op.backward_ast_gpu.input_fields[0].name, op.backward_ast_gpu.output_fields[0].name
        if len(fk.fields) != 2 or len(bk.fields) != 2:
            print('skip', name)
            continue
        cells = 1
        for s in shape:
            cells *= s
        esz = torch.empty((), dtype=dt).element_size()
        nbytes = cells * esz
        u, o, d, du = (torch.rand(shape, dtype=dt, device=dev) for _ in range(4))
        res = {'bytes_per_array': nbytes}
        # (a)+(b)
        res['fwd_alone_20'] = stats(per_launch(lambda i: fk(**{fin: u, fout: o}), 20))
        res['fwd_alone_400'] = stats(per_launch(lambda i: fk(**{fin: u, fout: o}), 400))
        res['pingpong_400'] = stats(per_launch(lambda i: fk(**{fin: (u, o)[i & 1], fout: (o, u)[i & 1]}), 400))

        def pair(i):
            if i & 1:
                bk(**{bin_: d, bout: du})
            else:
                fk(**{fin: u, fout: o})
        res['fwd_adj_alternating_400'] = stats(per_launch(pair, 400))
        res['fwd_adj_alternating_40'] = stats(per_launch(pair, 40))
        # (c) one allocation, out at (in + nbytes + delta)
        del o, d, du
        torch.cuda.empty_cache()
        big = torch.empty(2 * nbytes + (64 << 20), dtype=torch.uint8, device=dev)
        base = big.data_ptr()
        align = (-base) % (2 << 20)
        placement = {}
        for delta in (0, 256, 512, 1024, 2048, 4096, 8192, 16384, 65536, 1 << 18, 1 << 20, (1 << 20) + 4096, 3 << 19, 2 << 20,
                      (2 << 20) + 2048, 5 << 20, 32 << 20):
            a = big[align:align + nbytes].view(dt).view(shape)
            b0 = align + nbytes + delta
            b = big[b0:b0 + nbytes].view(dt).view(shape)
            a.copy_(u)
            ts = per_launch(lambda i: fk(**{fin: a, fout: b}), 30)
            placement[str(delta)] = round(stats(ts)['median'], 4)
        res['placement_out_minus_in_end'] = placement
        out[name] = res
        print(name, json.dumps(res), flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/r2_placement.json', 'w') as fh:
        json.dump(out, fh, indent=1)


if __name__ == '__main__':
    main()
