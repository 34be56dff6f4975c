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
