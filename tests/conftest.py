import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def make_data(n, d, seed, kind="lowrank", nq=0, rank=16, noise=0.3, normalize=True):
    """Synthetic embeddings. `lowrank`: x = normalise(z W + noise g), z in R^rank — neighbourhood
    structure comparable to real sentence embeddings; `iid`: isotropic Gaussian directions."""
    rng = np.random.default_rng(seed)
    if kind == "iid":
        f = lambda m: rng.standard_normal((m, d), dtype=np.float32)
    else:
        W = rng.standard_normal((rank, d), dtype=np.float32)
        f = lambda m: rng.standard_normal((m, rank), dtype=np.float32) @ W + noise * rng.standard_normal((m, d), dtype=np.float32)
    x = f(n)
    q = f(nq) if nq else None
    if normalize:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        if q is not None:
            q /= np.linalg.norm(q, axis=1, keepdims=True)
    return (np.ascontiguousarray(x, dtype=np.float32), None if q is None else np.ascontiguousarray(q, dtype=np.float32))


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def pkg():
    import leann_rs_b200
    return leann_rs_b200
