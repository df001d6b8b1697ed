import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        meta = json.load(f)
    arr = np.load(os.path.join(GOLDEN, "golden.npz"))
    return meta, arr


@pytest.fixture(scope="session")
def orc():
    from oracle import bind
    bind.build()
    return bind.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The real reference compiled from /root/reference (binary travels to the GPU box); None if absent."""
    from oracle import bind
    if not bind.ref_available("sse"):
        return None
    return bind.Ref("sse")


@pytest.fixture(scope="session")
def lib():
    import research_new_hnsw_b200 as pkg
    if not os.path.exists(pkg.lib_path()):
        pkg.build_library()
    return pkg


def gauss(seed, n, d):
    """Deterministic N(0,1) rows (numpy stream; used where the libstdc++ stream is not required)."""
    return np.random.default_rng(seed).standard_normal((n, d), dtype=np.float32)


@pytest.fixture(scope="session")
def graphs(orc, tmp_path_factory):
    """Reference-format index files built by the oracle (byte-identical to the reference's, see test_oracle.py),
    cached for the session: name -> dict(path, X, Q, metric, d, M)."""
    from oracle import bind
    root = tmp_path_factory.mktemp("graphs")
    specs = {
        "l2_d128": dict(metric=bind.L2, n=8000, d=128, M=16, efc=200),
        "l2_d100": dict(metric=bind.L2, n=4000, d=100, M=12, efc=100),
        "l2_d30": dict(metric=bind.L2, n=3000, d=30, M=8, efc=80),
        "ip_d96": dict(metric=bind.IP, n=6000, d=96, M=32, efc=120),
        "ip_d17": dict(metric=bind.IP, n=2000, d=17, M=6, efc=50),
        "lowrank_d128": dict(metric=bind.L2, n=10000, d=128, M=32, efc=100, lowrank=True),
    }
    out = {}
    for name, s in specs.items():
        if s.get("lowrank"):
            X = bind.lowrank_data(s["n"], s["d"], seed=1)
            Q = bind.lowrank_data(500, s["d"], seed=2)
        else:
            X = gauss(11, s["n"], s["d"])
            Q = gauss(12, 500, s["d"])
        if s["metric"] == bind.IP:
            X = X / np.linalg.norm(X, axis=1, keepdims=True)
        idx = orc.hnsw_new(s["metric"], s["d"], s["n"], s["M"], s["efc"])
        idx.add(X)
        path = str(root / (name + ".bin"))
        idx.save(path)
        out[name] = dict(path=path, X=X, Q=Q, **s)
    return out
