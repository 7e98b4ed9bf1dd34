"""adaface_b200 -- B200-native (sm_100a) implementation of AdaFace's data-parallel hot path.

Host-side mirrors of the reference's operator surface over the C-ABI library
``csrc/libadaface_b200.so`` (include/adaface_b200.h):

    AttnProcessor_LoRA_Capture, Attention, LoraDoraLinear, gen_gradient_scaler   (attn_processor.py)
    CrossAttention, FeedForward, BasicTransformerBlock, SpatialTransformer       (ldm_attention.py)
    ResBlock, Upsample, Downsample, LoraDoraConv2d                               (ldm_unet_blocks.py)
    UNetModel, TimestepEmbedSequential                                           (ldm_unet.py)
    DDIMSampler, UNetDenoiser                                                    (ddim.py)
    SubjBasisGenerator, Arc2FaceID2ImgPrompt, CLIPTextModelWrapper, CLIPAttentionMKV   (subj_basis_generator.py)

The directory is named ``adaface-dev_b200`` (not importable as is); import it as ``adaface_dev_b200``
(the alias package at the repository root).
"""
from . import _lib, ops  # noqa: F401
from .attn_processor import (AttnProcessor_LoRA_Capture, Attention, LoraDoraLinear, ScaleGrad, GradientScaler,  # noqa: F401
                             gen_gradient_scaler, img_mask_to_key_mask)
from .ldm_attention import CrossAttention, FeedForward, GEGLU, BasicTransformerBlock, SpatialTransformer  # noqa: F401
from .ldm_unet_blocks import ResBlock, Upsample, Downsample, LoraDoraConv2d  # noqa: F401
from .ldm_unet import UNetModel, TimestepEmbedSequential  # noqa: F401
from .subj_basis_generator import (SubjBasisGenerator, Arc2FaceID2ImgPrompt, FrozenCLIPTextEncoder, CLIPTextModelWrapper,  # noqa: F401
                                   CLIPAttentionMKV, CLIPTextConfig, template_ids)
from .ddim import DDIMSampler, UNetDenoiser, ddim_cfg_step, make_linear_alphas_cumprod  # noqa: F401
from .build import build  # noqa: F401
from .graphs import graphed, graphed_step, invalidate_trainable_packs  # noqa: F401
from . import parallel  # noqa: F401

__version__ = "0.1.0"
