"""Builds experiment variants of the remap translation unit into variants/libvkit_<name>.so
(same ABI; selected at run time with VKB_LIB=...).

    python tools/build_variants.py name1:-DFOO=1,-DBAR=2 name2: ...

Every variant compiles csrc/remap.cu with -DVKB_REMAP_ONLY_RGB (one instantiation) plus its own
flags and links it with the objects of the regular build (csrc/build/*.o).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vkit_b200 import build as vb  # noqa: E402


def main():
    vb.build()
    out_dir = os.path.join(ROOT, 'variants')
    os.makedirs(out_dir, exist_ok=True)
    nvcc = vb._find_nvcc()
    others = [vb._object_of(s) for s in vb.sources() if not s.endswith('remap.cu')]
    remap = os.path.join(vb.CSRC, 'remap.cu')

    def one(spec):
        name, _, flags = spec.partition(':')
        flags = [f for f in flags.split(',') if f]
        obj = os.path.join(out_dir, f'remap_{name}.o')
        lib = os.path.join(out_dir, f'libvkit_{name}.so')
        cmd = [nvcc] + vb.NVCC_FLAGS + ['-Xptxas', '-v', '-DVKB_REMAP_ONLY_RGB'] + flags + ['-c', remap, '-o', obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode:
            return name, proc.stderr
        info = [l for l in proc.stderr.split('\n') if 'grid_remap_tiles' in l or 'spill' in l or 'registers' in l]
        # keep the lines of the tiles kernel only
        text, keep = [], False
        for line in proc.stderr.split('\n'):
            if 'Compiling entry function' in line:
                keep = 'grid_remap_tiles' in line
            elif keep and ('spill' in line or 'registers' in line):
                text.append(line.strip())
        link = subprocess.run([nvcc, '--shared', '-gencode', 'arch=compute_100a,code=sm_100a', obj] + others
                              + ['-o', lib], capture_output=True, text=True)
        if link.returncode:
            return name, link.stderr
        return name, ' | '.join(text)

    with ThreadPoolExecutor(max_workers=8) as pool:
        for name, info in pool.map(one, sys.argv[1:]):
            print(f'{name}: {info}')


if __name__ == '__main__':
    main()
