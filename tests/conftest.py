import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_c1():
    return np.load(os.path.join(GOLDEN, "reference_c1.npz"))


@pytest.fixture(scope="session")
def golden_eval():
    return np.load(os.path.join(GOLDEN, "reference_eval.npz"))


def synthetic_case(n_users, n_items, d, seed):
    """Same generator as oracle/make_golden.py (legacy RandomState is stable across numpy versions)."""
    rs = np.random.RandomState(seed)
    content = rs.standard_normal((n_users + n_items, d)).astype(np.float32)
    collab = rs.standard_normal((n_users + n_items, d)).astype(np.float32)
    return content, collab


def synthetic_edges(n_users, n_items, seed, val_per_user=3, min_train=5, max_train=40):
    rs = np.random.RandomState(seed)
    train, val = [], []
    for u in range(n_users):
        deg = rs.randint(min_train, max_train + 1)
        items = rs.choice(n_items, size=deg + val_per_user, replace=False)
        ratings = rs.randint(1, 6, size=deg + val_per_user)
        for it, r in zip(items[:deg], ratings[:deg]):
            train.append((u, int(it), float(r)))
        if u % 7 != 0:
            for it, r in zip(items[deg:], ratings[deg:]):
                val.append((u, int(it), float(r)))
            if u % 5 == 0:
                val.append((u, int(items[0]), 5.0))
    return train, val


@pytest.fixture(scope="session")
def golden_lp():
    return np.load(os.path.join(GOLDEN, "reference_lp.npz"))


@pytest.fixture(scope="session")
def golden_c2():
    return np.load(os.path.join(GOLDEN, "reference_c2.npz"))


@pytest.fixture(scope="session")
def golden_c3():
    return np.load(os.path.join(GOLDEN, "reference_c3.npz"))


@pytest.fixture(scope="session")
def golden_gcn():
    return np.load(os.path.join(GOLDEN, "reference_gcn.npz"))


def gcn_case(g, ci):
    """Arguments of gcn_infer for case `ci` of reference_gcn.npz (oracle/make_golden_gcn.py)."""
    p = "c%d_" % ci
    n, C, F, L = [int(x) for x in g[p + "shape"]]
    nbr = [(g[p + "nbr_ptr%d" % l], g[p + "nbr_idx%d" % l]) for l in range(L)]
    names = ("node_emb", "content", "proj_w", "proj_b", "ln_g", "ln_b", "fc0_w", "fc0_b", "fc1_w", "fc1_b")
    return dict(shape=(n, C, F, L), nbr=nbr, previous=g[p + "previous_before"], previous_after=g[p + "previous_after"],
                h=g[p + "h"], **{k: g[p + k] for k in names})


@pytest.fixture(scope="session")
def golden_r2():
    return np.load(os.path.join(GOLDEN, "reference_r2.npz"))


def posneg_lists(u, n_items):
    """Positive / negative item ids of anchor user u in reference_c1.npz's and reference_r2.npz's pos/neg cases."""
    return [(u * 3 + j) % n_items for j in range(3)], [(u * 5 + j + 1) % n_items for j in range(2)]
