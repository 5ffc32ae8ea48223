#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: stall-reason totals and the hottest SASS instructions of a launch."""
import csv
import sys


def main(path, block_index=0, top=40):
    rows = list(csv.reader(open(path)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    b = blocks[block_index]
    hdr, data = b['rows'][0], b['rows'][1:]
    si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[si]) for r in data if len(r) > si and r[si].isdigit())
    inst = sum(int(r[ii]) for r in data if len(r) > ii and r[ii].isdigit())
    print(f'{b["name"]}: {len(data)} SASS lines, {tot} samples, {inst} warp instructions executed')
    agg = {h: 0 for _, h in stall_cols}
    for r in data:
        for i, h in stall_cols:
            if len(r) > i and r[i].isdigit():
                agg[h] += int(r[i])
    for h, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        if v:
            print(f'  {h:28s} {100.0 * v / max(tot, 1):6.2f} %')
    print('hottest instructions (samples %, executed, address, SASS, top stall):')
    hot = sorted((r for r in data if len(r) > si and r[si].isdigit()), key=lambda r: -int(r[si]))[:top]
    for r in hot:
        reasons = sorted(((int(r[i]), h) for i, h in stall_cols if len(r) > i and r[i].isdigit() and int(r[i])), reverse=True)[:2]
        print(f'  {100.0 * int(r[si]) / max(tot, 1):5.2f}  {r[ii]:>10s}  {r[0]:>6s}  {r[1][:70]:70s} {reasons}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
