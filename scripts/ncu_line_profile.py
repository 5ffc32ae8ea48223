#!/usr/bin/env python
"""Instructions executed / stall samples per CUDA source line of one launch of an ncu report: joins the SASS addresses of
`ncu --page source --csv` with the line table of `nvdisasm -g -c` on the cubin of the object that was profiled.

    python scripts/ncu_line_profile.py report.ncu-rep <launch index> <cubin> <mangled kernel substring> [top]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, idx, cubin, kname = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True).stdout.decode('utf-8', 'replace').split('\n')
    line_of, cur, inside = {}, ('?', 0), False
    for ln in dis:
        if ln.startswith('.text.'):
            inside = kname in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True).stdout.decode('utf-8', 'replace')
    blocks, curb = [], None
    for r in csv.reader(io.StringIO(out)):
        if r and r[0] == 'Kernel Name':
            curb = []
            blocks.append(curb)
        elif curb is not None:
            curb.append(r)
    h, data = blocks[idx][0], blocks[idx][1:]
    ai, ii, si = h.index('Address'), h.index('Instructions Executed'), h.index('# Samples')
    base = min(int(r[ai], 16) for r in data if r[ai].startswith('0x'))
    inst, smp = collections.Counter(), collections.Counter()
    for r in data:
        if not r[ai].startswith('0x') or not r[ii].isdigit():
            continue
        key = line_of.get(int(r[ai], 16) - base, ('?', 0))
        inst[key] += int(r[ii])
        smp[key] += int(r[si]) if r[si].isdigit() else 0
    ti, ts = sum(inst.values()), sum(smp.values())
    per_file = collections.Counter()
    for (f, _), n in inst.items():
        per_file[f] += n
    print('per file:', [(f, round(100.0 * n / ti, 1)) for f, n in per_file.most_common()])
    for key, n in inst.most_common(top):
        print(f'{key[0]:22s} {key[1]:5d}  {100.0 * n / ti:5.1f}% inst  {100.0 * smp[key] / max(ts, 1):5.1f}% samples')


if __name__ == '__main__':
    main()
