import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def state_dict():
    from said_b200.synth import synthetic_state_dict

    return synthetic_state_dict(seed=0)


_MODELS = {}


@pytest.fixture(scope="session")
def gpu_model(state_dict):
    """factory: prediction_type -> SAID_UNet1D on cuda:0 with the synthetic weights (cached)."""
    from said_b200.model.diffusion import SAID_UNet1D

    def make(prediction_type="epsilon"):
        if prediction_type not in _MODELS:
            m = SAID_UNet1D(prediction_type=prediction_type)
            m.load_state_dict(state_dict)
            m.to("cuda:0").eval()
            _MODELS[prediction_type] = m
        return _MODELS[prediction_type]

    return make
