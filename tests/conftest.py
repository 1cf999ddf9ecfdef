import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    names = (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return sorted(n for n in names if not n.startswith("refdump_"))   # refdump_*: N2 dump fixtures (test_gpu_api)


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    g["name"] = name
    return g


def golden_sticks(g):
    return {k: g[k] for k in ("xa", "ya", "xb", "yb", "xc", "yc", "length", "kind", "orig")}


def golden_element(g):
    return "fermidirac" if "Fermi" in str(g["element"]) else "linexp"


FULLNET_GATES = (("0", None, 0.0), ("_back", "back", 10.0), ("_total", "total", 10.0), ("_partial", "partial", 10.0))


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from networksim_cntfet_b200.engine import Engine
    return Engine()
