import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build the CUDA library and the oracle once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()
    return True


def make_setup(vehicle="iris", mode="traj", enu=True, **overrides):
    """(cfg, blob, model) for one of the shipped YAML configs."""
    from sde4mbrl_px4_b200 import config, model_io

    cfgd = config.load_yaml(os.path.join(ROOT, "configs", f"{vehicle}_{mode}.yaml"))
    width = overrides.pop("width", None)
    cfg = config.build_config(cfgd, convert_to_enu=enu, **overrides)
    model = model_io.synthetic_model(vehicle, width=width)
    return cfg, model.to_blob(), model


def random_states(n, seed=0):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 13), np.float64)
    x[:, 0:3] = rng.uniform(-1, 1, (n, 3)) + np.array([0, 0, 1.5])
    x[:, 3:6] = rng.normal(0, 0.3, (n, 3))
    q = np.array([1.0, 0, 0, 0]) + rng.normal(0, 0.1, (n, 4))
    x[:, 6:10] = q / np.linalg.norm(q, axis=1, keepdims=True)
    x[:, 10:13] = rng.normal(0, 0.2, (n, 3))
    return x.astype(np.float32)
