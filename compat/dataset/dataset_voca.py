"""The one thing the inference scripts take from script/dataset/dataset_voca.py: the CSV column order."""
from said_b200.util.blendshape import DEFAULT_BLENDSHAPE_CLASSES


class BlendVOCADataset:
    fps = 60
    default_blendshape_classes = list(DEFAULT_BLENDSHAPE_CLASSES)
