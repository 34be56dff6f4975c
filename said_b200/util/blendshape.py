"""Blendshape-coefficient I/O with the reference's interface (``said/util/blendshape.py:36-84``)."""
from __future__ import annotations

from typing import List

import numpy as np
import torch

# ARKit blendshape order of the CSV header (reference script/dataset/dataset_voca.py:99-132,
# data/ARKit_blendshapes.txt)
DEFAULT_BLENDSHAPE_CLASSES = [
    "jawForward", "jawLeft", "jawRight", "jawOpen", "mouthClose", "mouthFunnel", "mouthPucker", "mouthLeft",
    "mouthRight", "mouthSmileLeft", "mouthSmileRight", "mouthFrownLeft", "mouthFrownRight", "mouthDimpleLeft",
    "mouthDimpleRight", "mouthStretchLeft", "mouthStretchRight", "mouthRollLower", "mouthRollUpper",
    "mouthShrugLower", "mouthShrugUpper", "mouthPressLeft", "mouthPressRight", "mouthLowerDownLeft",
    "mouthLowerDownRight", "mouthUpperUpLeft", "mouthUpperUpRight", "cheekPuff", "cheekSquintLeft",
    "cheekSquintRight", "noseSneerLeft", "noseSneerRight",
]


def load_blendshape_coeffs(coeffs_path: str) -> torch.FloatTensor:
    """(T_b, num_classes) coefficients from a CSV with a header row (reference ``blendshape.py:36-51``)."""
    import pandas as pd

    df = pd.read_csv(coeffs_path)
    return torch.FloatTensor(df.values)


def save_blendshape_coeffs(coeffs: np.ndarray, classes: List[str], output_path: str) -> None:
    """CSV with one named column per class (reference ``blendshape.py:54-69``)."""
    import pandas as pd

    pout = pd.DataFrame(coeffs, columns=classes)
    pout.to_csv(output_path, index=False)


def save_blendshape_coeffs_image(coeffs: np.ndarray, output_path: str) -> None:
    """Grey-scale image, one row per class (reference ``blendshape.py:72-84``)."""
    from PIL import Image

    orig = (255 * coeffs.transpose()).round()
    img = Image.fromarray(orig).convert("L")
    img.save(output_path)


def save_blendshape_coeffs_batch(coeffs: np.ndarray, classes: List[str], output_paths: List[str]) -> None:
    """Write one CSV per clip of a (B, T_b, num_classes) result, byte-identical to ``save_blendshape_coeffs`` on each clip
    (the reference's ``script/test_inference.py:188-202`` goes through a pandas DataFrame per file: at batch 512 that host
    loop is the visible tail of an inference call).  Values are formatted once for the whole batch with
    the shortest round-trip float32 representation pandas uses."""
    coeffs = np.asarray(coeffs)
    if coeffs.ndim != 3 or coeffs.shape[0] != len(output_paths) or coeffs.shape[2] != len(classes):
        raise ValueError(
            f"coeffs must be (len(output_paths)={len(output_paths)}, T_b, len(classes)={len(classes)}); got {coeffs.shape}"
        )
    header = ",".join(classes)
    flat = coeffs.reshape(-1)
    # str() of a numpy float32 scalar is the shortest string that round-trips in float32 (what DataFrame.to_csv writes)
    cells = [str(v) for v in flat] if flat.dtype != np.float64 else [repr(float(v)) for v in flat]
    per_clip = coeffs.shape[1] * coeffs.shape[2]
    ncol = coeffs.shape[2]
    for b, path in enumerate(output_paths):
        clip = cells[b * per_clip:(b + 1) * per_clip]
        rows = [",".join(clip[r * ncol:(r + 1) * ncol]) for r in range(coeffs.shape[1])]
        with open(path, "w", newline="") as f:
            f.write(header + "\n" + "\n".join(rows) + ("\n" if rows else ""))
