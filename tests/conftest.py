import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def fold0(a):
    """-0.0 -> +0.0 (the reference yields signed zeros in IoU tables)."""
    return np.asarray(a) + np.float32(0.0)


def c4_inputs():
    """Seeded inputs of BASELINE config 1 -- the same draws as tests/golden/make_golden.py:c4_inputs (SEED + 31)."""
    rng = np.random.default_rng(2019 + 31)
    cls = rng.normal(0, 1, (12, 38, 64)).astype(np.float32)
    reg = rng.normal(0, 0.3, (48, 38, 64)).astype(np.float32)
    feat = rng.standard_normal((1, 1024, 38, 64), dtype=np.float32)
    cls_out = rng.normal(0, 2.5, (300, 21)).astype(np.float32)
    cls_out[:, 0] += 2.0
    reg_out = rng.normal(0, 0.5, (300, 84)).astype(np.float32)
    return cls, reg, feat, cls_out, reg_out


@pytest.fixture
def setknob(monkeypatch):
    """Set B2D_* development knobs for one test: the library reads its knobs once, so the environment change is
    followed by b2d_reload_knobs(); both are undone at teardown."""
    from b200det import _C

    def _set(**kv):
        for k, v in kv.items():
            monkeypatch.setenv(k, str(v))
        _C.reload_knobs()

    yield _set
    monkeypatch.undo()
    _C.reload_knobs()
