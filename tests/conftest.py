import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, 'tests')):
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on a B200)')


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests fail loudly without a device; plain runs skip them only when deselected.
    pass


@pytest.fixture(scope='session')
def hostsim():
    """g++ build of the shared per-element numerics (tests/hostsim/hostsim.cpp)."""
    import ctypes
    src = os.path.join(ROOT, 'tests', 'hostsim', 'hostsim.cpp')
    out = os.path.join(ROOT, 'tests', 'hostsim', 'libhostsim.so')
    deps = [src] + [os.path.join(ROOT, 'vkit_b200', 'csrc', f)
                    for f in ('vkb_math.cuh', 'vkb_lattice.cuh', 'vkb_color.cuh')]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(['g++', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-std=c++17',
                               src, '-o', out])
    return ctypes.CDLL(out)
