import torch
from torch import nn


class LearnedPerceptualImagePatchSimilarity(nn.Module):
    """No pretrained backbone is available offline: constructing it is allowed, evaluating it is not."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, preds, target):
        raise RuntimeError("stub LPIPS: no backbone weights in this image")
