"""Golden vectors for the evaluation embedder: the reference's OWN ``said/model/vae.py::BCVAE`` (imported unchanged from
/root/reference, run here) on seeded synthetic weights and synthetic coefficient windows.

    python tests/golden/make_golden_bcvae.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import eval_oracle as E  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_vae", "/root/reference/said/model/vae.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    vae = ref.BCVAE().eval()
    sd = E.synthetic_bcvae_state_dict(0)
    missing = vae.load_state_dict(sd, strict=False)
    assert all(not k.startswith("encoder.") or "num_batches_tracked" in k or "fc_logvar" in k for k in missing.missing_keys), missing
    g = torch.Generator().manual_seed(5)
    t = torch.arange(300, dtype=torch.float32)[None, :, None] / 60.0
    coeffs = (0.3 * (1 + torch.sin(6.2832 * (0.4 + 2 * torch.rand(6, 1, 32, generator=g)) * t + 6.2832 * torch.rand(6, 1, 32, generator=g)))).clamp(0, 0.75)
    step = 30
    nw = (300 - 120) // step + 1
    with torch.no_grad():
        lat_ref = torch.cat([vae.encode(coeffs[b:b + 1, s * step: s * step + 120]).mean for b in range(6) for s in range(nw)])
        lat_or = E.window_latents(sd, coeffs, step)
    d = float((lat_ref - lat_or).abs().max())
    print("oracle vs reference BCVAE.encode:", d, lat_ref.shape)
    assert d < 1e-5
    np.savez_compressed(os.path.join(HERE, "bcvae_windows.npz"), coeffs=coeffs.numpy(), step=np.int64(step), latents=lat_ref.numpy(),
                        oracle_vs_ref=np.float64(d))


if __name__ == "__main__":
    main()
