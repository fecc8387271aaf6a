"""Synthetic stand-in for a diffusers ``UNet2DConditionModel`` (benchmarks, smoke tests): random
weights of a named architecture behind the attributes `compile_unet` reads (`config`,
`state_dict()`, `device`, `dtype`, `forward`).  No checkpoints exist offline."""
import torch

from .unet_spec import param_shapes, random_state_dict, spec_from_config

SD15 = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                      "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
    layers_per_block=2, attention_head_dim=8, cross_attention_dim=768, norm_num_groups=32,
    norm_eps=1e-5, use_linear_projection=False, flip_sin_to_cos=True, freq_shift=0,
    transformer_layers_per_block=1, sample_size=64)

SDXL = dict(
    in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280),
    down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
    up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
    layers_per_block=2, attention_head_dim=(5, 10, 20), cross_attention_dim=2048,
    norm_num_groups=32, norm_eps=1e-5, use_linear_projection=True, flip_sin_to_cos=True,
    freq_shift=0, transformer_layers_per_block=(1, 2, 10), addition_embed_type="text_time",
    addition_time_embed_dim=256, projection_class_embeddings_input_dim=2816, sample_size=128)

TINY = dict(SD15, block_out_channels=(64, 128, 256, 256), attention_head_dim=(2, 2, 4, 4),
            cross_attention_dim=128, sample_size=32)

SVD = dict(
    in_channels=8, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
    down_block_types=("CrossAttnDownBlockSpatioTemporal",) * 3 + ("DownBlockSpatioTemporal",),
    up_block_types=("UpBlockSpatioTemporal",) + ("CrossAttnUpBlockSpatioTemporal",) * 3,
    layers_per_block=2, num_attention_heads=(5, 10, 20, 20), cross_attention_dim=1024,
    transformer_layers_per_block=1, addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=768, num_frames=25, sample_size=96)

SD_VAE = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
              layers_per_block=2, norm_num_groups=32, act_fn="silu", scaling_factor=0.18215, sample_size=512)

CONFIGS = {"sd15": SD15, "sdxl": SDXL, "tiny": TINY, "svd": SVD}


class SyntheticVAE:
    """Random-weight stand-in for a diffusers AutoencoderKL (decoder side) for `compile_vae`."""

    def __init__(self, config=None, seed=0, dtype=torch.float16, device="cuda"):
        import math
        from .vae_plan import vae_decoder_param_shapes, vae_spec_from_config
        self.config = dict(config or SD_VAE)
        self.dtype, self.device = dtype, torch.device(device)
        g = torch.Generator().manual_seed(seed)
        sd = {}
        for name, shape in vae_decoder_param_shapes(vae_spec_from_config(self.config)).items():
            if "norm" in name:
                t = torch.randn(shape, generator=g) * 0.3 + (1.0 if name.endswith("weight") else 0.0)
            else:
                fan = math.prod(shape[1:]) if len(shape) > 1 else shape[0]
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(max(fan, 1))
            sd[name] = t.to(dtype=dtype, device=self.device)
        self._sd = sd

    def state_dict(self):
        return self._sd

    def decode(self, *a, **k):
        raise RuntimeError("SyntheticVAE has no eager decode; compile it first")


class SyntheticUNet:
    def __init__(self, config, state_dict=None, seed=0, dtype=torch.float16, device="cuda"):
        self.config = dict(config)
        self.dtype = dtype
        self.device = torch.device(device)
        self.spec = spec_from_config(self.config)
        if state_dict is None:
            state_dict = random_state_dict(self.spec, seed=seed, dtype=dtype, device=self.device)
        self._sd = state_dict

    def state_dict(self):
        return self._sd

    def param_shapes(self):
        return param_shapes(self.spec)

    def forward(self, *a, **k):
        raise RuntimeError("SyntheticUNet has no eager forward; compile it first")

    def __call__(self, *a, **k):
        return self.forward(*a, **k)
