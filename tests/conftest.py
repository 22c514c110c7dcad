import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)          # tests/floatref.py, tests/srslte_ctypes.py


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def gpu():
    """(srsue_b200 module, Context) -- the product library; fails loudly when it is not built."""
    import torch
    import srsue_b200 as sg
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    ctx = sg.Context(0)
    yield sg, ctx
    ctx.close()
