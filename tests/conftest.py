import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def load_golden(name):
    """Golden case -> dict with the CSR matrix rebuilt as scipy CSR under 'X'."""
    import scipy.sparse as sp
    z = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    z["X"] = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
    z["n_bins"] = int(z["n_bins"])
    if "graph_data" in z:
        n = int(z["shape"][0])
        z["graph"] = sp.csr_matrix((z["graph_data"], z["graph_indices"], z["graph_indptr"]), shape=(n, n))
    return z


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from genevector_b200.engine import Engine
    return Engine()
