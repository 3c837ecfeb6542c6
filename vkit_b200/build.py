"""Builds vkit_b200/csrc/libvkit_b200.so for sm_100a with nvcc (in-tree, no JIT cache).

    python -m vkit_b200.build [--force] [--verbose]

Every `csrc/*.cu` is compiled to its own object (in parallel, only when it or a header changed)
and the objects are linked into the shared library.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
OBJ_DIR = os.path.join(CSRC, 'build')
LIB = os.path.join(CSRC, 'libvkit_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC',
]


def _find_nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _headers():
    return glob.glob(os.path.join(CSRC, '*.cuh')) + [
        os.path.join(_HERE, '..', 'include', 'vkit_b200.h')]


def _object_of(src):
    return os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    mtime = os.path.getmtime(target)
    return any(os.path.getmtime(d) > mtime for d in deps)


def needs_build():
    return _stale(LIB, sources() + _headers())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = _find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = _headers()
    extra = ['-Xptxas', '-v'] if verbose else []

    def compile_one(src):
        obj = _object_of(src)
        if not force and not _stale(obj, [src] + headers):
            return None
        cmd = [nvcc] + NVCC_FLAGS + extra + ['-c', src, '-o', obj]
        return cmd, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as pool:
        results = list(pool.map(compile_one, sources()))
    for res in results:
        if res is None:
            continue
        cmd, proc = res
        if verbose or proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
        if proc.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [nvcc, '--shared', '-gencode', 'arch=compute_100a,code=sm_100a'] + [
        _object_of(s) for s in sources()] + ['-o', LIB]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError('nvcc link failed: ' + ' '.join(cmd))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
