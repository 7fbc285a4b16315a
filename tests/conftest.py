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
def oracle():
    from oracle import native, reference_np
    native.build()
    return reference_np


@pytest.fixture(scope="session")
def s3fd_anchors_np(oracle):
    from dan_b200 import synthetic
    enc = oracle.AnchorEncoder(0.4, 0.4, [0.1, 0.1, 0.2, 0.2])
    return synthetic.build_anchors(enc, synthetic.pyramid_config("s3fd"))


@pytest.fixture(scope="session")
def dan_anchors_np(oracle):
    from dan_b200 import synthetic
    enc = oracle.AnchorEncoder(0.35, 0.35, [0.1, 0.1, 0.2, 0.2])
    return synthetic.build_anchors(enc, synthetic.pyramid_config("dan"))


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dan_b200  # noqa: F401  (fails loudly if libdan_b200.so is missing)
    from dan_b200 import _lib
    _lib.lib()
    return torch.device("cuda", 0)


def to_dev(arr, device, dtype=None):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(device)
