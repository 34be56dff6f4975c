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


LARGE_FAMILY_CONFIG = dict(   # tests/golden/make_golden_large.py: a reduced member of the wav2vec2-large family
    hidden_size=256, num_hidden_layers=3, num_attention_heads=4, intermediate_size=512,
    feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True,
    num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16, vocab_size=32,
)


@pytest.fixture(scope="session")
def large_family():
    """(Wav2Vec2Config, {"audio_encoder.*": tensor}) regenerated from the seed through transformers' own Wav2Vec2Model
    (the third-party module the reference subclasses); the golden file's weights_abs_sum guards the regeneration."""
    import numpy as np
    from transformers import Wav2Vec2Config, Wav2Vec2Model

    cfg = Wav2Vec2Config(**LARGE_FAMILY_CONFIG)
    torch.manual_seed(0)
    hf = Wav2Vec2Model(cfg).eval()
    sd = {"audio_encoder." + k: v.detach().clone() for k, v in hf.state_dict().items()}
    gd = np.load(os.path.join(GOLDEN, "audio_encoder_large_family_1s.npz"))
    got = sum(float(v.double().abs().sum()) for v in sd.values())
    if abs(got - float(gd["weights_abs_sum"])) > 1e-6 * abs(got):
        pytest.skip("transformers initialises Wav2Vec2Model differently here: the large-family golden cannot be regenerated")
    return cfg, sd
