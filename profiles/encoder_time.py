"""Audio encoder (Wav2Vec2 + interpolation) for 64 x 5 s clips: milliseconds per call in each encoder precision mode, and the
deviation of the tensor-core modes from the fp32 kernels.

    python profiles/encoder_time.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from said_b200.model.diffusion import SAID_UNet1D  # noqa: E402
from said_b200.synth import synthetic_batch, synthetic_state_dict  # noqa: E402

m = SAID_UNet1D()
m.load_state_dict(synthetic_state_dict(0))
m.to("cuda:0").eval()
wave = synthetic_batch(64, 5.0).to("cuda:0")
ref = None
for mode in ("fp32", "fp16x3", "tf32x3"):
    m.encoder_precision = mode
    with torch.no_grad():
        for _ in range(2):
            emb = m.get_audio_embedding(wave, 300)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            emb = m.get_audio_embedding(wave, 300)
        e1.record()
        torch.cuda.synchronize()
    if ref is None:
        ref = emb.clone()
    print(f"encoder {mode}: {e0.elapsed_time(e1) / 3:.2f} ms per 64 x 5 s clips; max |emb - fp32 emb| = {float((emb - ref).abs().max()):.3e}")
