"""ORACLE (test infrastructure, NOT product code).

PyTorch restatement of diffusers ``UNetSpatioTemporalConditionModel`` (Stable Video Diffusion),
the module stable-fast compiles for SVD pipelines
(/root/reference/src/sfast/compilers/diffusion_pipeline_compiler.py:101-103,
/root/reference/examples/optimize_stable_video_diffusion_pipeline.py): BASELINE.json configs[3].

PARITY STATUS: **unpinned.**  The arithmetic lives in the un-vendored dependency ``diffusers``
(absent from this image, no network), the reference handles the model generically by tracing and
holds no SVD test or vector.  The module tree, parameter names and forward semantics below follow
diffusers 0.24-0.27 (`models/unets/unet_spatio_temporal_condition.py`, `unet_3d_blocks.py`,
`transformers/transformer_temporal.py`, `attention.py: TemporalBasicTransformerBlock`,
`resnet.py: SpatioTemporalResBlock / TemporalResnetBlock / AlphaBlender`), including its
batch-interleaved temporal cross-attention context (`time_context` is laid out [H*W, B] while the
temporal tokens are laid out [B, H*W]; kept as published).  Structural check: the parameter total
of the SVD-XT configuration is asserted in tests/test_oracle.py.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference legs may
import this module.
"""
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet_oracle import (Attention, BasicTransformerBlock, Downsample2D, FeedForward, ResnetBlock2D,
                          TimestepEmbedding, Upsample2D, timestep_embedding)


@dataclass
class SVDConfig:
    in_channels: int = 8
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlockSpatioTemporal",) * 3 + ("DownBlockSpatioTemporal",)
    up_block_types: Tuple[str, ...] = ("UpBlockSpatioTemporal",) + ("CrossAttnUpBlockSpatioTemporal",) * 3
    layers_per_block: int = 2
    num_attention_heads: Tuple[int, ...] = (5, 10, 20, 20)
    transformer_layers_per_block: int = 1
    cross_attention_dim: int = 1024
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 768
    num_frames: int = 25
    norm_num_groups: int = 32
    sample_size: int = 96

    def get(self, k, d=None):
        return getattr(self, k, d)


def svd_xt_config() -> SVDConfig:
    return SVDConfig()


def svd_tiny_config() -> SVDConfig:
    """Same topology at 1/5 width, 6 frames: seconds-scale parity tests."""
    return SVDConfig(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4),
                     cross_attention_dim=128, addition_time_embed_dim=32,
                     projection_class_embeddings_input_dim=96, num_frames=6, sample_size=32)


class AlphaBlender(nn.Module):
    """merge_strategy="learned_with_images" with image_only_indicator == 0 everywhere (what the
    SVD UNet passes): alpha = sigmoid(mix_factor)."""

    def __init__(self, alpha: float):
        super().__init__()
        self.mix_factor = nn.Parameter(torch.tensor([alpha]))

    def forward(self, x_spatial, x_temporal):
        alpha = torch.sigmoid(self.mix_factor).to(x_spatial.dtype)
        return alpha * x_spatial + (1.0 - alpha) * x_temporal


class TemporalResnetBlock(nn.Module):
    def __init__(self, c, temb_dim, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, c, eps=eps)
        self.conv1 = nn.Conv3d(c, c, (3, 1, 1), padding=(1, 0, 0))
        self.time_emb_proj = nn.Linear(temb_dim, c)
        self.norm2 = nn.GroupNorm(groups, c, eps=eps)
        self.conv2 = nn.Conv3d(c, c, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, x, temb):  # x [B, C, F, H, W], temb [B, F, temb_dim]
        h = self.conv1(F.silu(self.norm1(x)))
        t = self.time_emb_proj(F.silu(temb))[:, :, :, None, None].permute(0, 2, 1, 3, 4)
        h = h + t
        h = self.conv2(F.silu(self.norm2(h)))
        return x + h


class SpatioTemporalResBlock(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups, eps):
        super().__init__()
        self.spatial_res_block = ResnetBlock2D(cin, cout, temb_dim, groups, eps)
        self.temporal_res_block = TemporalResnetBlock(cout, temb_dim, groups, eps)
        self.time_mixer = AlphaBlender(0.5)

    def forward(self, x, temb, num_frames):  # x [B*F, C, H, W], temb [B*F, temb_dim]
        x = self.spatial_res_block(x, temb)
        bf, c, h, w = x.shape
        b = bf // num_frames
        xs = x.reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        xt = self.temporal_res_block(xs, temb.reshape(b, num_frames, -1))
        y = self.time_mixer(xs, xt)
        return y.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


class TemporalBasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, ctx_dim):
        super().__init__()
        self.norm_in = nn.LayerNorm(dim)
        self.ff_in = FeedForward(dim)
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, ctx_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, num_frames, ctx):  # x [B*F, S, C]
        bf, s, c = x.shape
        b = bf // num_frames
        x = x.reshape(b, num_frames, s, c).permute(0, 2, 1, 3).reshape(b * s, num_frames, c)
        x = self.ff_in(self.norm_in(x)) + x
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), ctx) + x
        x = self.ff(self.norm3(x)) + x
        return x.reshape(b, s, num_frames, c).permute(0, 2, 1, 3).reshape(bf, s, c)


class TransformerSpatioTemporalModel(nn.Module):
    def __init__(self, dim, heads, ctx_dim, depth, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(dim, heads, ctx_dim) for _ in range(depth)])
        self.temporal_transformer_blocks = nn.ModuleList(
            [TemporalBasicTransformerBlock(dim, heads, ctx_dim) for _ in range(depth)])
        self.time_pos_embed = TimestepEmbedding(dim, dim * 4, out_dim=dim)
        self.time_mixer = AlphaBlender(0.5)
        self.proj_out = nn.Linear(dim, dim)
        self.dim = dim

    def forward(self, x, ctx, num_frames):  # x [B*F, C, H, W], ctx [B*F, 1, ctx_dim]
        bf, c, hh, ww = x.shape
        b = bf // num_frames
        # diffusers: first frame's context, broadcast as [H*W, B] (sic) then flattened
        first = ctx.reshape(b, num_frames, -1, ctx.shape[-1])[:, 0]
        time_ctx = first[None, :].expand(hh * ww, b, first.shape[1], first.shape[2])
        time_ctx = time_ctx.reshape(hh * ww * b, first.shape[1], first.shape[2])
        res = x
        h = self.norm(x).permute(0, 2, 3, 1).reshape(bf, hh * ww, c)
        h = self.proj_in(h)
        frames = torch.arange(num_frames, device=x.device).repeat(b, 1).reshape(-1)
        t_emb = timestep_embedding(frames, self.dim, True, 0).to(h.dtype)
        emb = self.time_pos_embed(t_emb)[:, None, :]
        for blk, tblk in zip(self.transformer_blocks, self.temporal_transformer_blocks):
            h = blk(h, ctx)
            h_mix = tblk(h + emb, num_frames, time_ctx)
            h = self.time_mixer(h, h_mix)
        h = self.proj_out(h)
        return h.reshape(bf, hh, ww, c).permute(0, 3, 1, 2) + res


class STDownBlock(nn.Module):
    def __init__(self, cfg, cin, cout, temb_dim, has_attn, heads, eps, add_down):
        super().__init__()
        n, g = cfg.layers_per_block, cfg.norm_num_groups
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(cin if i == 0 else cout, cout, temb_dim, g, eps)
                                      for i in range(n)])
        self.attentions = nn.ModuleList([
            TransformerSpatioTemporalModel(cout, heads, cfg.cross_attention_dim,
                                           cfg.transformer_layers_per_block, g)
            for _ in range(n)]) if has_attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ctx, nf):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb, nf)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx, nf)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class STMidBlock(nn.Module):
    def __init__(self, cfg, c, temb_dim, heads):
        super().__init__()
        g = cfg.norm_num_groups
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(c, c, temb_dim, g, 1e-5) for _ in range(2)])
        self.attentions = nn.ModuleList([TransformerSpatioTemporalModel(
            c, heads, cfg.cross_attention_dim, cfg.transformer_layers_per_block, g)])

    def forward(self, x, temb, ctx, nf):
        x = self.resnets[0](x, temb, nf)
        x = self.attentions[0](x, ctx, nf)
        return self.resnets[1](x, temb, nf)


class STUpBlock(nn.Module):
    def __init__(self, cfg, cin, cout, cprev, temb_dim, has_attn, heads, add_up):
        super().__init__()
        n, g = cfg.layers_per_block + 1, cfg.norm_num_groups
        res = []
        for i in range(n):
            skip = cin if i == n - 1 else cout
            rin = cprev if i == 0 else cout
            res.append(SpatioTemporalResBlock(rin + skip, cout, temb_dim, g, 1e-6))
        self.resnets = nn.ModuleList(res)
        self.attentions = nn.ModuleList([
            TransformerSpatioTemporalModel(cout, heads, cfg.cross_attention_dim,
                                           cfg.transformer_layers_per_block, g)
            for _ in range(n)]) if has_attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, ctx, nf):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb, nf)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx, nf)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


@dataclass
class SVDOutput:
    sample: torch.Tensor


class UNetSpatioTemporalConditionModel(nn.Module):
    def __init__(self, cfg: SVDConfig):
        super().__init__()
        self.config = cfg
        boc = cfg.block_out_channels
        temb_dim = boc[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_dim)
        self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, temb_dim)
        nb = len(boc)
        self.down_blocks = nn.ModuleList()
        cout = boc[0]
        for i, t in enumerate(cfg.down_block_types):
            cin, cout = cout, boc[i]
            attn = t.startswith("CrossAttn")
            self.down_blocks.append(STDownBlock(cfg, cin, cout, temb_dim, attn, cfg.num_attention_heads[i],
                                                1e-6 if attn else 1e-5, i != nb - 1))
        self.mid_block = STMidBlock(cfg, boc[-1], temb_dim, cfg.num_attention_heads[-1])
        self.up_blocks = nn.ModuleList()
        rboc, rheads = list(reversed(boc)), list(reversed(cfg.num_attention_heads))
        cout = rboc[0]
        for i, t in enumerate(cfg.up_block_types):
            cprev, cout = cout, rboc[i]
            cin = rboc[min(i + 1, nb - 1)]
            self.up_blocks.append(STUpBlock(cfg, cin, cout, cprev, temb_dim, t.startswith("CrossAttn"),
                                            rheads[i], i != nb - 1))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids, return_dict=True):
        cfg = self.config
        b, nf = sample.shape[:2]
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], device=sample.device)
        t = timestep.reshape(-1).to(sample.device)
        t = t.expand(b) if t.numel() == 1 else t
        emb = self.time_embedding(timestep_embedding(t, cfg.block_out_channels[0], True, 0).to(sample.dtype))
        tid = timestep_embedding(added_time_ids.flatten(), cfg.addition_time_embed_dim, True, 0)
        emb = emb + self.add_embedding(tid.reshape(b, -1).to(sample.dtype))
        x = sample.flatten(0, 1)
        emb = emb.repeat_interleave(nf, dim=0)
        ctx = encoder_hidden_states.repeat_interleave(nf, dim=0)
        x = self.conv_in(x)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, ctx, nf)
            skips.extend(outs)
        x = self.mid_block(x, emb, ctx, nf)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, ctx, nf)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        x = x.reshape(b, nf, *x.shape[1:])
        if not return_dict:
            return (x,)
        return SVDOutput(sample=x)


def build_svd_unet(cfg: SVDConfig, seed: int = 0, dtype=torch.float32, device="cpu", randomize=True):
    """Seeded default-init model; norm affines and the AlphaBlender mix factors are randomised so
    that parity tests see them (PyTorch's defaults 1 / 0 / 0.5 would hide wiring mistakes)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = UNetSpatioTemporalConditionModel(cfg)
    if randomize:
        with torch.no_grad():
            for mod in m.modules():
                if isinstance(mod, (nn.GroupNorm, nn.LayerNorm)):
                    mod.weight.copy_(1.0 + 0.3 * torch.randn_like(mod.weight))
                    mod.bias.copy_(0.3 * torch.randn_like(mod.bias))
                elif isinstance(mod, AlphaBlender):
                    mod.mix_factor.copy_(torch.randn(1) * 1.5)
    torch.random.set_rng_state(g)
    return m.to(device=device, dtype=dtype).eval()
