"""Turn gpurun_out/*.ncu-rep and the launch-list CSV into the committed summaries under profiles/.

    python scripts/summarize_ncu.py r1     (reads gpurun_out/r1_*; writes profiles/r1_*.md, profiles/traffic.json)
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
GP = os.path.join(ROOT, 'gpurun_out')

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def raw_metrics(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
    os.makedirs(OUT, exist_ok=True)
    traffic = {}
    tpath = os.path.join(OUT, 'traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    # ---- launch list
    lpath = os.path.join(GP, '%s_launches_c3.csv' % tag)
    if os.path.exists(lpath):
        lines = [l for l in open(lpath) if l.startswith('"')]
        rows = list(csv.DictReader(io.StringIO(''.join(lines))))
        total = sum(float(r['Metric Value']) for r in rows)
        by = {}
        for r in rows:
            k = r['Kernel Name'].split('(')[0][:90]
            d = by.setdefault(k, [0, 0.0, r['Grid Size'], r['Block Size']])
            d[0] += 1
            d[1] += float(r['Metric Value'])
        with open(os.path.join(OUT, '%s_launches_c3.md' % tag), 'w') as fh:
            cmd = ('python bench.py --steps 2 --warmup 3 --only-headline --no-cpu-baseline --e2e-steps 1' if tag != 'r1' else
                   'python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline')
            fh.write('# ncu launch list — `%s` (C3, 1 B200)\n\n' % cmd)
            fh.write('`ncu --metrics gpu__time_duration.sum --clock-control none -c 400`; per-launch times are cold-cache and '
                     'serialised: compare SHARES. %d launches, %.3f ms of kernel time.\n\n' % (len(rows), total / 1e6))
            fh.write('| kernel | launches | total ms | share | avg ms | grid | block |\n|---|---|---|---|---|---|---|\n')
            for k, d in sorted(by.items(), key=lambda kv: -kv[1][1]):
                fh.write('| `%s` | %d | %.3f | %.1f %% | %.3f | %s | %s |\n' % (k, d[0], d[1] / 1e6, 100 * d[1] / total,
                                                                             d[1] / d[0] / 1e6, d[2], d[3]))
            fh.write('\nThe `psad_*` kernels are this repo\'s (NVRTC, sm_100a): the whole-field forward / adjoint launches of the '
                     'first, warm-up and timed steps (~1.4 ms per launch), the fused forward+adjoint and fused-pair diagnostics, '
                     'and the per-chunk launches of the host-streamed e2e passes. Everything else is torch\'s fill / RNG / copy '
                     'kernels used to create the synthetic inputs and to stage the e2e comparison, outside the timed region.\n')
        import shutil
        shutil.copy(lpath, os.path.join(OUT, '%s_launches_c3.csv' % tag))
    # ---- full captures
    names = {'c3_fwd': ('c3', 'forward'), 'c2_fwd': ('c2', 'forward'), 'c4_fwd': ('c4', 'forward'), 'c5_fwd': ('c5', 'forward'),
             'c5_bwd': ('c5', 'adjoint'), 'c3_x2e': ('c3', 'fused_pair')}
    with open(os.path.join(OUT, '%s_ncu_full_summary.md' % tag), 'w') as fh:
        fh.write('# ncu --set full summaries (%s)\n\nOne launch per kernel, captured with `ncu --set full --clock-control none '
                 '--import-source on` under the bench command line of each workload.\n' % tag)
        for key, (wl, which) in names.items():
            rep = os.path.join(GP, '%s_%s.ncu-rep' % (tag, key))
            if not os.path.exists(rep):
                continue
            m = raw_metrics(rep)
            fh.write('\n## %s %s kernel\n\n| metric | value | unit |\n|---|---|---|\n' % (wl.upper(), which))
            for w in WANT:
                if w in m:
                    fh.write('| %s | %s | %s |\n' % (w, m[w][0], m[w][1]))
            def gb(name):
                v, u = m[name]
                return float(v) * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[u]
            tot = gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum')
            traffic.setdefault(wl, {})['%s_dram_bytes_per_launch' % which] = tot
            dur, du = m['gpu__time_duration.sum']
            dur = float(dur) * {'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1}[du]
            fh.write('\nDRAM traffic %.3f GB per launch, %.0f GB/s under the profiler.\n' % (tot / 1e9, tot / dur / 1e9))
            stalls = sorted(((float(v[0]), k) for k, v in m.items() if 'smsp__average_warps_issue_stalled' in k
                             and k.endswith('per_issue_active.ratio')), reverse=True)[:6]
            fh.write('\nTop stall reasons (warps per issue-active cycle): ' +
                     ', '.join('%s %.2f' % (k.split('stalled_')[1].split('_per_')[0], v) for v, k in stalls) + '\n')
    json.dump(traffic, open(tpath, 'w'), indent=1, sort_keys=True)
    print('wrote', os.listdir(OUT))


if __name__ == '__main__':
    main()
