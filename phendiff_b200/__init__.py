"""phendiff_b200 — B200-native implementation of PhenDiff's class-conditional DDIM inversion + regeneration path.

Public surface mirrors the reference's (src/__init__.py:1-24, for the hot path only):
CustomCondUNet2DModel, ConditionalDDIMPipeline, CustomEmbedding, DDIMScheduler / DDIMInverseScheduler, `_inversion`, `_ddib`, `_classifier_free_guidance_forward_start`, `_linear_interp_custom_guidance_inverted_start`;
the training step lives in `phendiff_b200.training`.
"""
from .cond_unet_2d import CustomCondUNet2DModel, UNet2DOutput
from .custom_embedding import CustomEmbedding
from .pipeline_conditional_ddim import ConditionalDDIMPipeline, ImagePipelineOutput
from .schedulers import DDIMInverseScheduler, DDIMScheduler
from .utils_img2img import (_classifier_free_guidance_forward_start, _custom_guided_generation, _ddib, _inversion,
                            _linear_interp_custom_guidance_inverted_start, ddib_transfer)
from ._lib import PhenDiffB200Error

__all__ = ["CustomCondUNet2DModel", "UNet2DOutput", "CustomEmbedding", "ConditionalDDIMPipeline", "ImagePipelineOutput",
           "DDIMScheduler", "DDIMInverseScheduler", "_ddib", "_inversion", "ddib_transfer",
           "_classifier_free_guidance_forward_start", "_custom_guided_generation",
           "_linear_interp_custom_guidance_inverted_start", "PhenDiffB200Error"]
