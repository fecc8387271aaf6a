"""``sfast.compilers.diffusion_pipeline_compiler`` -- same names, fields and call convention as
/root/reference/src/sfast/compilers/diffusion_pipeline_compiler.py:20-190, with the UNet hot path
executed by the B200-native runtime (``sfast_b200``).

    from sfast.compilers.diffusion_pipeline_compiler import compile, CompilationConfig
    config = CompilationConfig.Default()
    config.enable_cuda_graph = True
    pipe = compile(pipe, config)          # same pipeline object, unet.forward replaced

Field mapping (all 11 reference fields are accepted):
  enable_cuda_graph      -> capture the whole UNet step as one CUDA graph per (B, H, W, dtype)
  enable_jit, enable_jit_freeze, enable_cnn_optimization, enable_fused_linear_geglu,
  prefer_lowp_gemm, enable_xformers, enable_triton, memory_format
                         -> accepted; they select nothing: the native path always runs NHWC with
                            every fusion on (fp32 accumulation, erf-GELU)
  preserve_parameters    -> True (default): in-place parameter updates (LoRA switch, reference
                            README.md:228-265) are picked up by the next call -- the packed copies
                            are refreshed in place when the parameters' version counters move;
                            False: weights are frozen at first call (``rebind()`` refreshes)
  trace_scheduler        -> accepted, no-op (scheduler is outside the hot path)
"""
import logging
from dataclasses import dataclass

import torch

from sfast.utils import gpu_device

logger = logging.getLogger()


class CompilationConfig:

    @dataclass
    class Default:
        memory_format: torch.memory_format = (
            torch.channels_last if gpu_device.device_has_tensor_core() else
            torch.contiguous_format)
        enable_jit: bool = True
        enable_jit_freeze: bool = True
        preserve_parameters: bool = True
        enable_cnn_optimization: bool = gpu_device.device_has_tensor_core()
        enable_fused_linear_geglu: bool = gpu_device.device_has_capability(8, 0)
        prefer_lowp_gemm: bool = True
        enable_xformers: bool = False
        enable_cuda_graph: bool = False
        enable_triton: bool = False
        trace_scheduler: bool = False


def compile(m, config):
    """Compile a diffusers pipeline in place and return it (reference `compile`, :81-124)."""
    m.unet = compile_unet(m.unet, config)
    if getattr(m, 'controlnet', None) is not None:
        logger.warning('sfast (B200 build): controlnet is left on its eager path '
                       '(outside the UNet hot path, see DESIGN.md)')
    if getattr(m, 'vae', None) is not None:
        m.vae = compile_vae(m.vae, config)
    # text encoders (reference :93-112: lazy trace + CUDA graph of text_encoder / text_encoder_2)
    for attr in ('text_encoder', 'text_encoder_2'):
        if getattr(m, attr, None) is not None:
            setattr(m, attr, compile_text_encoder(getattr(m, attr), config))
    # SVD's CLIP vision tower (reference :100-103)
    if getattr(m, 'image_encoder', None) is not None:
        m.image_encoder = compile_image_encoder(m.image_encoder, config)
    if getattr(config, 'trace_scheduler', False):
        logger.warning('sfast (B200 build): trace_scheduler is ignored (scheduler.step stays eager)')
    return m


def compile_image_encoder(m, config):
    """Replace ``m.forward`` of a transformers CLIPVisionModel / CLIPVisionModelWithProjection (the SVD
    pipeline's image_encoder, reference :100-103) with the B200-native tower; anything else is returned
    unchanged with a warning."""
    from sfast_b200.runtime import compile_image_encoder_module, require_b200
    cfg = getattr(m, 'config', None)
    if cfg is None or not hasattr(m, 'vision_model') or getattr(cfg, 'model_type', 'clip_vision_model') != 'clip_vision_model':
        logger.warning('sfast (B200 build): %s is not a CLIP vision model; left on its eager path', type(m).__name__)
        return m
    device = m.device if hasattr(m, 'device') else torch.device(
        'cuda' if torch.cuda.is_available() else 'cpu')
    require_b200(torch.device(device))
    return compile_image_encoder_module(m, enable_cuda_graph=bool(config.enable_cuda_graph))


def compile_text_encoder(m, config):
    """Replace ``m.forward`` of a transformers CLIPTextModel / CLIPTextModelWithProjection with the
    B200-native encoder (reference :93-112).  Any other encoder class, or one on a non-sm_100
    device, is returned unchanged with a warning."""
    from sfast_b200.runtime import compile_text_encoder_module, require_b200
    cfg = getattr(m, 'config', None)
    if cfg is None or not hasattr(m, 'text_model') or getattr(cfg, 'model_type', 'clip_text_model') != 'clip_text_model':
        logger.warning('sfast (B200 build): %s is not a CLIP text model; left on its eager path', type(m).__name__)
        return m
    device = m.device if hasattr(m, 'device') else torch.device(
        'cuda' if torch.cuda.is_available() else 'cpu')
    require_b200(torch.device(device))
    return compile_text_encoder_module(m, enable_cuda_graph=bool(config.enable_cuda_graph))


def compile_unet(m, config):
    """Replace ``m.forward`` with the B200-native UNet step (reference `compile_unet`, :127-151).

    There is no CPU / eager fallback: the module must live on an sm_100 CUDA device."""
    from sfast_b200.runtime import compile_unet_module, require_b200
    device = m.device if hasattr(m, 'device') else torch.device(
        'cuda' if torch.cuda.is_available() else 'cpu')
    require_b200(torch.device(device))
    return compile_unet_module(m, enable_cuda_graph=bool(config.enable_cuda_graph),
                               preserve_parameters=bool(getattr(config, 'preserve_parameters', True)))


def compile_vae(m, config):
    """Replace ``m.decode`` with the B200-native VAE decoder (reference `compile_vae`, :154-190, which
    wraps the sub-modules in auto-trace hooks); `encode` stays eager.  A VAE without a `decode`
    method, or one on a non-sm_100 device, is returned unchanged with a warning (the reference, too,
    only hooks what it can trace)."""
    from sfast_b200.runtime import compile_vae_module, require_b200
    if not hasattr(m, 'decode') or not hasattr(m, 'config'):
        logger.warning('sfast (B200 build): VAE without decode()/config is left on its eager path')
        return m
    device = m.device if hasattr(m, 'device') else torch.device(
        'cuda' if torch.cuda.is_available() else 'cpu')
    require_b200(torch.device(device))
    return compile_vae_module(m, enable_cuda_graph=bool(config.enable_cuda_graph))
