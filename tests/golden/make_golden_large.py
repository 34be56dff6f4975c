"""Golden vector for the wav2vec2-large FAMILY of audio encoders (``feat_extract_norm="layer"``,
``conv_bias=True``, ``do_stable_layer_norm=True``), produced by the UNMODIFIED reference module
``said/model/wav2vec2.py::ModifiedWav2Vec2Model`` (run here, where /root/reference is mounted).

The configuration is a reduced member of the family (hidden 256, 4 heads, 3 layers, ffn 512; the real large model is
1024 / 16 / 24 / 4096 with the same code path) so that the weights can be REGENERATED in the tests from the seed through
``transformers.Wav2Vec2Model(config)`` instead of being committed; ``weights_abs_sum`` guards that regeneration.

    python tests/golden/make_golden_large.py
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import said_oracle as O  # noqa: E402
from said_b200.synth import normalise_waveform, synthetic_waveform  # noqa: E402

LARGE_FAMILY_CONFIG = dict(
    hidden_size=256, num_hidden_layers=3, num_attention_heads=4, intermediate_size=512,
    feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True,
    num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16, vocab_size=32,
)
# The real wav2vec2-large (facebook/wav2vec2-large-960h-lv60 architecture): 1024 / 16 heads / 24 layers / 4096, 315 M parameters.
LARGE_FULL_CONFIG = dict(
    hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
    feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True,
    num_conv_pos_embeddings=128, num_conv_pos_embedding_groups=16, vocab_size=32,
)


def main():
    from transformers import Wav2Vec2Config

    said = types.ModuleType("said")
    said.__path__ = [REF + "/said"]
    sys.modules["said"] = said
    model = types.ModuleType("said.model")
    model.__path__ = [REF + "/said/model"]
    sys.modules["said.model"] = model
    ref = importlib.import_module("said.model.wav2vec2")

    for conf, fname in ((LARGE_FAMILY_CONFIG, "audio_encoder_large_family_1s.npz"), (LARGE_FULL_CONFIG, "audio_encoder_large_full_1s.npz")):
        make_one(ref, Wav2Vec2Config(**conf), fname)


def make_one(ref, cfg, fname):
    torch.manual_seed(0)
    m = ref.ModifiedWav2Vec2Model(cfg).eval()
    sd = {"audio_encoder." + k: v.detach().clone() for k, v in m.state_dict().items()}
    wave = torch.from_numpy(normalise_waveform(synthetic_waveform(0, 1.0)))[None]
    with torch.no_grad():
        emb = m(wave, num_frames=60).last_hidden_state
        emb_o = O.wav2vec2_forward(sd, wave, 60, stable_layer_norm=True)
        emb_o64 = O.wav2vec2_forward({k: v.double() for k, v in sd.items()}, wave.double(), 60, stable_layer_norm=True)
    d_or = float((emb - emb_o).abs().max())
    d_64 = float((emb.double() - emb_o64).abs().max())
    print("oracle vs reference:", d_or, " fp32 vs fp64 floor:", d_64, " mean |emb|:", float(emb.abs().mean()))
    assert d_or < 1e-4
    np.savez_compressed(
        os.path.join(HERE, fname), wave=wave.numpy(), emb=emb.numpy(),
        emb64=emb_o64.float().numpy(), weights_abs_sum=np.float64(sum(float(v.double().abs().sum()) for v in sd.values())),
        oracle_vs_ref=np.float64(d_or), fp32_vs_fp64=np.float64(d_64))


if __name__ == "__main__":
    main()
