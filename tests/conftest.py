import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def pytest_collection_modifyitems(config, items):
    # GPU tests must never silently pass on a machine without a GPU: they are skipped, loudly, and the driver runs
    # them with `-m gpu` on a B200 where skipping is an error (see test_gpu_*.py::test_cuda_is_present).
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "head_small_g8.npz")))


@pytest.fixture(scope="session")
def golden_resnet():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "resnet_g32.npz")))


def head_params_from(d):
    return {k: d[k] for k in ("in_proj_weight", "in_proj_bias", "out_proj_weight", "out_proj_bias",
                              "classifier_weight", "classifier_bias")}
