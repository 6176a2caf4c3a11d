"""Per-launch kernel time next to nvidia-smi clocks / power / temperature while one kernel runs back to back: is the drift of
a long run (scripts/placement_bench.py: C4 1.13 -> 1.31 ms over 400 launches) a clock, power or memory effect?

    python scripts/sustain_trace.py c4 600
"""
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pystencils_autodiff_b200.backends._torch_native import CompiledKernel, numpy_dtype_to_torch
from pystencils_autodiff_b200.configs import CONFIG_SHAPES, make_config

Q = ('clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory,clocks_event_reasons.sw_power_cap,'
     'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown')


def main():
    name, iters = sys.argv[1], int(sys.argv[2])
    shape = CONFIG_SHAPES[name]['shape']
    op = make_config(name, shape=shape, boundary_handling='zeros')
    tun = {}
    for kv in os.environ.get('PSAD_TUNE', '').split(','):
        if kv:
            a, b = kv.split('=')
            tun[a] = bool(int(b)) if a in ('carry', 'shuffle', 'plane_sums', 'arrival', 'linopt', 'cross_cse', 'lds_pair') else int(b)
    from pystencils_autodiff_b200.emit import MarchTuning
    k = CompiledKernel(op.forward_ast_gpu, MarchTuning(**tun) if tun else None)
    dt = numpy_dtype_to_torch(k.fields[0].dtype.numpy_dtype)
    arrs = {f.name: torch.rand(shape, dtype=dt, device='cuda') for f in k.fields}
    for _ in range(3):
        k(**arrs)
    torch.cuda.synchronize()
    time.sleep(3.0)                      # start from an idle GPU
    lines = []
    proc = subprocess.Popen(['nvidia-smi', '-i', '0', '--query-gpu=' + Q, '--format=csv,noheader,nounits', '-lms', '20'],
                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

    def pump():
        for ln in proc.stdout:
            lines.append((time.perf_counter(), ln.strip()))
    threading.Thread(target=pump, daemon=True).start()
    time.sleep(0.3)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    t0 = time.perf_counter()
    evs[0].record()
    for i in range(iters):
        k(**arrs)
        evs[i + 1].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    time.sleep(0.2)
    proc.terminate()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(iters)]
    print('%s: %d launches in %.3f s' % (name, iters, t1 - t0))
    acc = 0.0
    for b in range(0, iters, 50):
        chunk = ts[b:b + 50]
        t_mid = t0 + (acc + sum(chunk) / 2) * 1e-3
        acc += sum(chunk)
        near = min(lines, key=lambda x: abs(x[0] - t_mid))[1] if lines else ''
        print('launch %4d-%4d: mean %.4f ms  min %.4f  | sm,mem MHz, W, T_gpu, T_mem, pcap, sw_th, hw, hw_th = %s'
              % (b, b + len(chunk) - 1, sum(chunk) / len(chunk), min(chunk), near))


if __name__ == '__main__':
    main()
