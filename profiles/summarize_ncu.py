#!/usr/bin/env python
"""Turn an .ncu-rep (one kernel launch, `ncu --set full --import-source on`) into the text summary committed
under profiles/:  key launch metrics, stall reasons, executed-instruction mix, hottest SASS instructions.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/r01_<name>.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__block_size',
        'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def page(rep, name, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'] + list(extra), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    raw = page(rep, 'raw')
    hdr, units, vals = raw[0], raw[1], raw[2]
    d = dict(zip(hdr, zip(units, vals)))
    print('# ncu summary of %s' % rep)
    print('kernel: %s' % d.get('Kernel Name', ('', '?'))[1])
    for k in KEYS:
        if k in d:
            print('%-70s %-16s %s' % (k, d[k][0], d[k][1]))
    print('\n## warp stall reasons (per issued instruction)')
    st = [(h, float(v[1])) for h, v in d.items() if h.startswith('smsp__average_warps_issue_stalled_')
          and h.endswith('_per_issue_active.ratio')]
    for h, v in sorted(st, key=lambda x: -x[1])[:8]:
        print('%-40s %.3f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
    src = page(rep, 'source', ['--print-source', 'sass'])
    if len(src) > 3:
        h = src[1]
        ix = {n: i for i, n in enumerate(h)}
        data = src[2:]

        def f(r, k):
            try:
                return float(r[ix[k]])
            except Exception:
                return 0.0
        tot = sum(f(r, 'Instructions Executed') for r in data)
        mix = collections.Counter()
        for r in data:
            op = r[ix['Source']].strip().split()
            o = op[1] if op[0].startswith('@') else op[0]
            mix[o.split('.')[0]] += f(r, 'Instructions Executed')
        print('\n## executed warp-instruction mix (%d SASS instructions, %.3g executed)' % (len(data), tot))
        for k, v in mix.most_common(14):
            print('%-10s %6.2f %%' % (k, 100 * v / tot))
        fp64 = sum(v for k, v in mix.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU'))
        print('FP64-pipe share of issued instructions: %.1f %%' % (100 * fp64 / tot))
        print('\n## hottest instructions by stall samples')
        for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:12]:
            print('%8d  %s' % (f(r, '# Samples'), r[ix['Source']].strip()[:80]))


if __name__ == '__main__':
    main(sys.argv[1])
