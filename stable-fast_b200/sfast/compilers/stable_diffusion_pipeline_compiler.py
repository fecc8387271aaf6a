"""Deprecated alias kept by the reference
(/root/reference/src/sfast/compilers/stable_diffusion_pipeline_compiler.py:1-8)."""
from .diffusion_pipeline_compiler import *  # noqa: F401,F403
from .diffusion_pipeline_compiler import CompilationConfig, compile, compile_unet, compile_vae  # noqa: F401
