"""B200-native runtime for the diffusion-UNet denoise step (C-ABI kernels + schedule builder)."""
