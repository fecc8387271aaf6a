"""ORACLE (test infrastructure, NOT product code).

PyTorch restatement of the decoder half of diffusers ``AutoencoderKL`` (the SD-1.5 / SDXL VAE), the
module `compile_vae` wraps in the reference
(/root/reference/src/sfast/compilers/diffusion_pipeline_compiler.py:154-190).  PARITY STATUS:
**unpinned** -- the arithmetic lives in the un-vendored dependency ``diffusers`` (absent here) and
the reference holds no VAE test or vector.  Module tree and parameter names follow diffusers
(`AutoencoderKL.post_quant_conv`, `decoder.conv_in`, `decoder.mid_block.{resnets,attentions}`,
`decoder.up_blocks.N.{resnets,upsamplers}`, `decoder.conv_norm_out`, `decoder.conv_out`);
structural check: the decoder + post_quant_conv parameter total of the SD VAE, 49,490,199
(tests/test_oracle.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.18215
    sample_size: int = 512

    def get(self, k, d=None):
        return getattr(self, k, d)


def sd_vae_config() -> VAEConfig:
    return VAEConfig()


def tiny_vae_config() -> VAEConfig:
    return VAEConfig(block_out_channels=(64, 64, 128, 128), sample_size=64)


class VaeResnet(nn.Module):
    """ResnetBlock2D without a time embedding (eps 1e-6)."""

    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (self.conv_shortcut(x) if self.conv_shortcut is not None else x) + h


class VaeAttention(nn.Module):
    """Single-head self-attention over the H*W tokens (diffusers Attention with group_norm,
    residual_connection=True, biased projections)."""

    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Identity()])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).reshape(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o)
        return o.transpose(1, 2).reshape(b, c, h, w) + x


class VaeMidBlock(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnet(c, c, groups), VaeResnet(c, c, groups)])
        self.attentions = nn.ModuleList([VaeAttention(c, groups)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class VaeUpsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class VaeUpBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnet(cin if i == 0 else cout, cout, groups) for i in range(n)])
        self.upsamplers = nn.ModuleList([VaeUpsample(cout)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Decoder(nn.Module):
    def __init__(self, cfg: VAEConfig):
        super().__init__()
        boc, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, boc[-1], 3, padding=1)
        self.mid_block = VaeMidBlock(boc[-1], g)
        self.up_blocks = nn.ModuleList()
        rboc = list(reversed(boc))
        cout = rboc[0]
        for i, c in enumerate(rboc):
            cin, cout = cout, c
            self.up_blocks.append(VaeUpBlock(cin, cout, cfg.layers_per_block + 1, g, i != len(boc) - 1))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            x = blk(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class AutoencoderKLDecoder(nn.Module):
    """`AutoencoderKL` restricted to what `decode()` touches: post_quant_conv + decoder."""

    def __init__(self, cfg: VAEConfig):
        super().__init__()
        self.config = cfg
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.decoder = Decoder(cfg)

    @property
    def dtype(self):
        return self.post_quant_conv.weight.dtype

    @property
    def device(self):
        return self.post_quant_conv.weight.device

    def decode(self, z, return_dict=True):
        x = self.decoder(self.post_quant_conv(z))
        if not return_dict:
            return (x,)
        return DecoderOutput(sample=x)


def build_vae(cfg: VAEConfig, seed: int = 0, dtype=torch.float32, device="cpu"):
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = AutoencoderKLDecoder(cfg)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, nn.GroupNorm):
                mod.weight.copy_(1.0 + 0.3 * torch.randn_like(mod.weight))
                mod.bias.copy_(0.3 * torch.randn_like(mod.bias))
    torch.random.set_rng_state(g)
    return m.to(device=device, dtype=dtype).eval()
