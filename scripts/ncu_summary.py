#!/usr/bin/env python
"""Text summary of an `ncu --set full --import-source on` report for profiles/: per-launch key metrics (raw page) and,
for one launch, the stall mix plus the hottest source lines (source page, needs -lineinfo).

    python scripts/ncu_summary.py report.ncu-rep [launch index for the source table] [header text] > profiles/x.txt
"""
import csv
import io
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
           'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
           'sm__cycles_elapsed.avg']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True).stdout.decode('utf-8', 'replace')
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    if len(sys.argv) > 3:
        print('# ' + sys.argv[3])
    raw = page(rep, 'raw')
    hdr, units, rows = raw[0], raw[1], raw[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print('Kernel Name', [r[col['Kernel Name']][:70] for r in rows])
    for m in METRICS:
        if m in col:
            print(m, [units[col[m]]] + [r[col[m]] for r in rows])
    src = page(rep, 'source')
    blocks, cur = [], None
    for r in src:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    if not blocks:
        return
    b = blocks[min(idx, len(blocks) - 1)]
    h, data = b['rows'][0], b['rows'][1:]
    si, ii = h.index('# Samples'), h.index('Instructions Executed')
    srci = h.index('Source') if 'Source' in h else 1
    stall = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    tot = sum(int(r[si]) for r in data if len(r) > si and r[si].isdigit())
    inst = sum(int(r[ii]) for r in data if len(r) > ii and r[ii].isdigit())
    print(f'source table of launch {idx} ({b["name"][:60]}): {inst} warp instructions, {tot} samples')
    agg = {}
    for r in data:
        for i, c in stall:
            if len(r) > i and r[i].isdigit():
                agg[c] = agg.get(c, 0) + int(r[i])
    print('stall mix:', [(c[6:], round(100.0 * v / max(tot, 1), 1)) for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]])
    hot = sorted((r for r in data if len(r) > si and r[si].isdigit()), key=lambda r: -int(r[si]))[:25]
    for r in hot:
        top = sorted(((int(r[i]), c[6:]) for i, c in stall if len(r) > i and r[i].isdigit() and int(r[i])), reverse=True)[:2]
        print(f'  {100.0 * int(r[ii]) / max(inst, 1):5.1f}% inst {100.0 * int(r[si]) / max(tot, 1):5.1f}% smp  {r[srci][:90]:90s} {top}')


if __name__ == '__main__':
    main()
