import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def state_dict():
    from fastvideocodec_b200.synthetic import init_state_dict
    return init_state_dict(seed=0)


@pytest.fixture(scope="session")
def golden_pframe_64():
    return load_golden("pframe_64.npz")


@pytest.fixture(scope="session")
def golden_pframe_128():
    return load_golden("pframe_128.npz")


@pytest.fixture(scope="session")
def golden_gop_64():
    return load_golden("gop_64.npz")


@pytest.fixture(scope="session")
def golden_ops():
    return load_golden("ops.npz")


@pytest.fixture(scope="session")
def real_state_dict(state_dict):
    """init_state_dict(0) with opticFlow.* replaced by the reference's own pretrained SpyNet weights
    (tests/golden/spynet_real.npz, written by oracle/gen_golden_r2.py from DVC/flow_pretrain_np; |w|max ~ 5)."""
    sd = dict(state_dict)
    real = load_golden("spynet_real.npz")
    assert set(real) == {k for k in sd if k.startswith("opticFlow.")}
    sd.update(real)
    return sd


@pytest.fixture(scope="session")
def golden_pframe_real_128():
    return load_golden("pframe_real_128.npz")


@pytest.fixture(scope="session")
def golden_pframe_L6_256():
    return load_golden("pframe_L6_256.npz")


@pytest.fixture(scope="session")
def golden_hd_gop10():
    return load_golden("hd_gop10.npz")
