"""CustomEmbedding — the reference's class-embedding wrapper (reference: src/custom_embedding/custom_embedding.py:6-17):
an `nn.Embedding(num_classes, class_embedding_dim)` exposed as `inner_module`, with the config protocol so it can be a
pipeline component.  It is only used by the reference's Stable-Diffusion path; the DDIM UNet embeds classes internally
(cond_unet_2d.py:146-147), where the lookup is fused into the embedding kernel.  The lookup here is a plain row gather.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .config import ConfigMixin


class CustomEmbedding(nn.Module, ConfigMixin):
    config_name = "config.json"

    def __init__(self, num_classes: int, class_embedding_dim: int):
        super().__init__()
        self.register_to_config(num_classes=num_classes, class_embedding_dim=class_embedding_dim)
        self.inner_module = nn.Embedding(num_classes, class_embedding_dim)

    @property
    def dtype(self):
        return self.inner_module.weight.dtype

    @property
    def device(self):
        return self.inner_module.weight.device

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.inner_module(x)
