import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def lib_built():
    """libmgb200.so, built on demand (nvcc cross-compiles without a GPU)."""
    from multigrid_dolfinx_b200 import build
    return build.build()


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    c, lc, lf, seed, mu, K = (int(x) for x in d["meta"])
    return d, dict(c=c, coarsest_level=lc, finest_level=lf, perm_seed=None if seed < 0 else seed, mu1=mu, mu2=mu), K


GOLDEN_CASES = ["cfg1_lex_mu2", "cfg1_lex_mu50", "cfg1_perm_mu2", "cfg1_perm_mu50", "proto_perm_mu50", "l4_perm_mu3"]
GOLDEN_SMALL = GOLDEN_CASES[:5]
