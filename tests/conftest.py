import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qp-zk-circuits_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

P = 0xFFFFFFFF00000001


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def splitmix_columns(seed, num_cols, n):
    """SURVEY.md 8d input recipe: SplitMix64 stream, rejection-sampled < p, filled column-major."""
    total = num_cols * n
    out = np.empty(0, np.uint64)
    state = np.uint64(seed)
    need = total
    chunks = []
    idx0 = 0
    with np.errstate(over="ignore"):
        while need > 0:
            m = int(need * 1.01) + 16
            k = np.arange(idx0 + 1, idx0 + m + 1, dtype=np.uint64)
            z = state + k * np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            z = z[z < np.uint64(P)]
            chunks.append(z[:need])
            need -= min(need, z.size)
            idx0 += m
    out = np.concatenate(chunks)
    return out.reshape(num_cols, n)


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def qpzk():
    import qpzk as q
    return q


@pytest.fixture(scope="session")
def ctx(qpzk):
    return qpzk.Context(0)
