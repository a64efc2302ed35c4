"""pytest configuration.  `gpu` marks tests that need a B200; everything else runs
on the CPU-only build container (`python -m pytest tests -m "not gpu"`)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


@pytest.fixture(scope='session')
def built():
    """Compile the engine + host-sim libraries (cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return g


@pytest.fixture(scope='session')
def hostsim(built):
    """Host build of the device numerical core (test infrastructure)."""
    return C.CDLL(built.HOSTSIM)


@pytest.fixture(scope='session')
def oracle():
    import pwinds_oracle
    return pwinds_oracle


@pytest.fixture(scope='session')
def sed(oracle):
    g = golden('solar_sed')
    return oracle.Sed(g['wavelength'], g['flux_lambda'])


def P(a):
    return a.ctypes.data_as(C.c_void_p)
