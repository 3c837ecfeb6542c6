"""Per-source-line executed-instruction profile of one kernel from an ncu report.

    python tools/region_profile.py <report.ncu-rep> <kernel regex> [min share %]

Uses `ncu --page source --print-source cuda,sass --csv` (the report must have been captured with
--import-source on and the library built with -lineinfo) and prints, in source order, the share
of executed warp instructions and stall samples per source line, plus an opcode histogram."""
import csv
import subprocess
import sys
from collections import Counter


def load(rep, kernel_regex):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass',
                          '--csv', '-k', 'regex:' + kernel_regex], capture_output=True, text=True).stdout
    lines, ops = [], Counter()
    path, header, seen = None, None, set()
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == 'File Path':
            path = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            continue
        if r[0] == 'Line No':
            header = {c: i for i, c in enumerate(r)}
            continue
        if header is None or len(r) < 8:
            continue
        ie, ss = header['Instructions Executed'], header['# Samples']
        if r[0]:  # a source line with its aggregate
            if r[ie].isdigit():
                lines.append((path, int(r[0]), r[1].strip(), int(r[ie]), int(r[ss]) if r[ss].isdigit() else 0))
        elif r[2].startswith('0x') and r[2] not in seen and r[ie].isdigit():
            seen.add(r[2])
            txt = r[3].strip()
            op = txt.split()[1] if txt.startswith('@') else txt.split()[0]
            base = op.split('.')[0]
            wide = [x for x in op.split('.')[1:] if x in ('U8', '64', '128', 'U16', '4A', '2A')]
            ops[base + ('.' + '.'.join(wide) if wide else '')] += int(r[ie])
    return lines, ops


if __name__ == '__main__':
    rep, kre = sys.argv[1:3]
    floor = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
    lines, ops = load(rep, kre)
    tot = sum(ops.values())
    samp = sum(l[4] for l in lines)
    print(f'total warp instructions {tot}, samples {samp}')
    for path, ln, code, n, s in lines:
        if 100.0 * n / tot >= floor or 100.0 * s / max(samp, 1) >= 1.0:
            print(f'{n / tot:6.2%} inst {s / max(samp, 1):6.2%} stall  {path}:{ln}  {code[:100]}')
    print('--- opcodes')
    for k, v in ops.most_common(30):
        print(f'{k:14s} {v:12d} {v / tot:6.2%}')
