"""Static SASS instruction count per source line of one kernel (needs -lineinfo).

    python tools/sass_lines.py <object or .so> <kernel name substring> [file substring]

Extracts the sm_100a cubin with cuobjdump, disassembles it with `nvdisasm -g -c` and attributes
every instruction to the most recent `//## File "...", line N` marker (inlined frames: the
innermost line).  For a rolled loop body the static count per line is the dynamic count per
iteration, which is how the remap kernel's per-pixel budgets in DESIGN.md were read.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, kernel = sys.argv[1], sys.argv[2]
    file_filter = sys.argv[3] if len(sys.argv) > 3 else ''
    work = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=work, check=True,
                   stdout=subprocess.DEVNULL)
    cubins = [os.path.join(work, f) for f in os.listdir(work) if f.endswith('.cubin')]
    text = ''
    for cubin in cubins:
        text += subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    lines = text.split('\n')
    start = None
    for i, line in enumerate(lines):
        if line.startswith('.text.') and kernel in line:
            start = i
            break
    if start is None:
        raise SystemExit('kernel not found')
    counts = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    cur = ('?', 0)
    total = 0
    for line in lines[start + 1:]:
        if line.startswith('.text.') or line.startswith('.section'):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m:
            counts[cur] += 1
            ops[cur][m.group(1)] += 1
            total += 1
    print('total instructions', total)
    for (fname, ln), n in sorted(counts.items()):
        if file_filter and file_filter not in fname:
            continue
        top = ' '.join(f'{k}:{v}' for k, v in ops[(fname, ln)].most_common(6))
        print(f'{fname}:{ln:5d} {n:5d}  {top}')


if __name__ == '__main__':
    main()
