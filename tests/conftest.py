import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def golden_inputs(g):
    """(x, edge_index int64, y) of a models_*.pt file (chameleon stores x as non-zero coordinates)."""
    if "x" in g:
        x = g["x"]
    else:
        x = torch.zeros(g["x_shape"])
        nz = g["x_nz"].long()
        x[nz[:, 0], nz[:, 1]] = 1.0
    return x, g["edge_index"].long(), g["y"]
