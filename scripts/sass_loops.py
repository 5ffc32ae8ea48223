#!/usr/bin/env python
"""Instruction-class histogram of the loops of a kernel's SASS (cuobjdump -sass output): offline estimate of the issue,
FMA-pipe, ALU-pipe and SFU cycles per loop iteration (pipe costs measured by scripts/ubench/pipes.cu on B200)."""
import re
import sys
from collections import Counter

FMA1 = ('FFMA', 'FADD', 'FMUL', 'IMAD', 'FSWZADD')
FMA2 = ('FFMA2', 'FADD2', 'FMUL2')
MUFU = ('MUFU',)
LSU = ('LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL', 'LDTM', 'ATOMS', 'RED', 'LDC', 'LDSM')


def classify(op):
    base = op.split('.')[0]
    if base in FMA2:
        return 'fma2'
    if base in FMA1:
        return 'fma'
    if base in MUFU:
        return 'mufu'
    if base in LSU:
        return 'lsu'
    if base in ('BRA', 'BSSY', 'BSYNC', 'EXIT', 'WARPSYNC', 'BAR', 'NOP', 'SYNCS', 'ELECT', 'VOTE', 'R2UR', 'S2R', 'UMOV', 'CS2R') \
            or base.startswith('U'):
        return 'ctl'
    return 'alu'


def main(path, min_len=60):
    ins = []
    for line in open(path):
        m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    addr_index = {a: i for i, (a, _, _) in enumerate(ins)}
    loops = []
    for i, (a, op, rest) in enumerate(ins):
        if op.startswith('BRA'):
            t = re.search(r'0x([0-9a-f]+)', rest)
            if t:
                ta = int(t.group(1), 16)
                if ta < a and ta in addr_index and i - addr_index[ta] >= min_len:
                    loops.append((addr_index[ta], i))
    for lo, hi in loops:
        body = ins[lo:hi + 1]
        c = Counter(classify(op) for _, op, _ in body)
        ops = Counter(op.split('.')[0] for _, op, _ in body)
        n = len(body)
        fma = c['fma'] + 2 * c['fma2']
        print(f'loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {n} instr | fma-pipe {fma} (fma {c["fma"]}, packed {c["fma2"]}) | '
              f'alu {c["alu"]} (x2 = {2 * c["alu"]} cyc) | mufu {c["mufu"]} (x8 = {8 * c["mufu"]} cyc) | lsu {c["lsu"]} | ctl {c["ctl"]}')
        print('   ', ', '.join(f'{k} {v}' for k, v in ops.most_common(24)))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
