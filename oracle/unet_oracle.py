"""ORACLE (test infrastructure, NOT product code).

CPU/fp32-capable PyTorch restatement of the arithmetic on the hot path that stable-fast
accelerates: the per-timestep ``UNet2DConditionModel.forward`` of a diffusers pipeline, which
the reference wraps at ``src/sfast/compilers/diffusion_pipeline_compiler.py:127-151``.

PARITY STATUS: **UNet-level parity unpinned.**  The UNet arithmetic is not in /root/reference;
it lives in the un-vendored, un-pinned dependency ``diffusers>=0.19.0``
(``/root/reference/setup.py:246-249``), which is absent from this image, and the reference's
own pipeline tests assert nothing about outputs
(``tests/compilers/test_stable_diffusion_pipeline_compiler.py:38-39,370,436``).  This file
restates the published diffusers algorithm (module tree, parameter names and shapes follow
diffusers ``UNet2DConditionModel`` for the SD-1.5 and SDXL-base configs) and is pinned
structurally by the exact parameter totals 859,520,964 (SD-1.5) and 2,567,463,684 (SDXL-base)
(``tests/test_oracle.py``).  The *fused-operator* semantics the reference itself defines are
pinned op-by-op against the reference's own operator tests in ``oracle/ops_oracle.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` leg may import this module.
"""
import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    """Subset of diffusers ``UNet2DConditionModel.config`` that the hot path reads."""
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    down_block_types: Tuple[str, ...] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D",
                                         "CrossAttnDownBlock2D", "DownBlock2D")
    up_block_types: Tuple[str, ...] = ("UpBlock2D", "CrossAttnUpBlock2D",
                                       "CrossAttnUpBlock2D", "CrossAttnUpBlock2D")
    layers_per_block: int = 2
    # diffusers' historical naming bug: for these configs `attention_head_dim` is the HEAD COUNT.
    attention_head_dim: Tuple[int, ...] = (8, 8, 8, 8)
    transformer_layers_per_block: Tuple[int, ...] = (1, 1, 1, 1)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    addition_embed_type: Optional[str] = None
    addition_time_embed_dim: Optional[int] = None
    projection_class_embeddings_input_dim: Optional[int] = None
    sample_size: int = 64

    def get(self, k, d=None):
        return getattr(self, k, d)


def sd15_config() -> UNetConfig:
    return UNetConfig()


def sdxl_config() -> UNetConfig:
    return UNetConfig(
        block_out_channels=(320, 640, 1280),
        down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D"),
        up_block_types=("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"),
        attention_head_dim=(5, 10, 20),
        transformer_layers_per_block=(1, 2, 10),
        cross_attention_dim=2048,
        use_linear_projection=True,
        addition_embed_type="text_time",
        addition_time_embed_dim=256,
        projection_class_embeddings_input_dim=2816,
        sample_size=128,
    )


def tiny_config() -> UNetConfig:
    """Same topology as SD-1.5 at 1/5 width: used for seconds-scale parity tests."""
    return UNetConfig(block_out_channels=(64, 128, 256, 256), attention_head_dim=(2, 2, 4, 4),
                      cross_attention_dim=128, sample_size=32)


def timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos: bool,
                       downscale_freq_shift: float, max_period: int = 10000) -> torch.Tensor:
    """diffusers ``get_timestep_embedding`` (sinusoidal), computed in fp32."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32,
                                                    device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, dim, out_dim=None):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, out_dim or dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        # reference fusions on this sub-graph: group_norm+silu
        # (/root/reference/src/sfast/jit/passes/triton_passes.py:68-88), conv+bias+add
        # (/root/reference/src/sfast/jit/passes/__init__.py:310-350).
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Attention(nn.Module):
    def __init__(self, dim, heads, ctx_dim=None):
        super().__init__()
        self.heads = heads
        ctx_dim = ctx_dim or dim
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, s, c = x.shape
        h = self.heads
        q = self.to_q(x).view(b, s, h, c // h).transpose(1, 2)
        k = self.to_k(ctx).view(b, ctx.shape[1], h, c // h).transpose(1, 2)
        v = self.to_v(ctx).view(b, ctx.shape[1], h, c // h).transpose(1, 2)
        # softmax(QK^T/sqrt(d))V: what xformers.memory_efficient_attention computes at
        # /root/reference/src/sfast/libs/xformers/xformers_attention.py:36-42
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(b, s, c)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        # hidden first, gate second: /root/reference/src/sfast/jit/passes/__init__.py:643-648
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, ctx_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, ctx_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), ctx) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, ctx_dim, depth, groups, linear_proj):
        super().__init__()
        self.linear_proj = linear_proj
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(dim, heads, ctx_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        b, c, hh, ww = x.shape
        res = x
        h = self.norm(x)
        if not self.linear_proj:
            h = self.proj_in(h)
            h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        else:
            h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
            h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        if not self.linear_proj:
            h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
            h = self.proj_out(h)
        else:
            h = self.proj_out(h)
            h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
        return h + res


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg, cin, cout, temb_dim, has_attn, heads, depth, add_down):
        super().__init__()
        n = cfg.layers_per_block
        self.resnets = nn.ModuleList([
            ResnetBlock2D(cin if i == 0 else cout, cout, temb_dim, cfg.norm_num_groups,
                          cfg.norm_eps) for i in range(n)])
        if has_attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(cout, heads, cfg.cross_attention_dim, depth,
                                   cfg.norm_num_groups, cfg.use_linear_projection)
                for _ in range(n)])
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, x, temb, ctx):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg, c, temb_dim, heads, depth):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(c, c, temb_dim, cfg.norm_num_groups, cfg.norm_eps) for _ in range(2)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(c, heads, cfg.cross_attention_dim, depth, cfg.norm_num_groups,
                               cfg.use_linear_projection)])

    def forward(self, x, temb, ctx):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ctx)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg, cin, cout, cprev, temb_dim, has_attn, heads, depth, add_up):
        super().__init__()
        n = cfg.layers_per_block + 1
        res = []
        for i in range(n):
            skip = cin if i == n - 1 else cout
            rin = cprev if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, temb_dim, cfg.norm_num_groups,
                                     cfg.norm_eps))
        self.resnets = nn.ModuleList(res)
        if has_attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(cout, heads, cfg.cross_attention_dim, depth,
                                   cfg.norm_num_groups, cfg.use_linear_projection)
                for _ in range(n)])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, x, skips, temb, ctx):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ctx)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


@dataclass
class UNetOutput:
    sample: torch.Tensor


class UNet2DConditionModel(nn.Module):
    """Restatement of diffusers ``UNet2DConditionModel`` (state_dict keys match diffusers')."""

    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.config = cfg
        boc = cfg.block_out_channels
        temb_dim = boc[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb_dim)
        if cfg.addition_embed_type == "text_time":
            self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim,
                                                   temb_dim)
        nb = len(boc)
        self.down_blocks = nn.ModuleList()
        cout = boc[0]
        for i, t in enumerate(cfg.down_block_types):
            cin, cout = cout, boc[i]
            self.down_blocks.append(DownBlock(
                cfg, cin, cout, temb_dim, t.startswith("CrossAttn"), cfg.attention_head_dim[i],
                cfg.transformer_layers_per_block[i], i != nb - 1))
        self.mid_block = MidBlock(cfg, boc[-1], temb_dim, cfg.attention_head_dim[-1],
                                  cfg.transformer_layers_per_block[-1])
        self.up_blocks = nn.ModuleList()
        rboc = list(reversed(boc))
        rheads = list(reversed(cfg.attention_head_dim))
        rdepth = list(reversed(cfg.transformer_layers_per_block))
        cout = rboc[0]
        for i, t in enumerate(cfg.up_block_types):
            cprev, cout = cout, rboc[i]
            cin = rboc[min(i + 1, nb - 1)]
            self.up_blocks.append(UpBlock(
                cfg, cin, cout, cprev, temb_dim, t.startswith("CrossAttn"), rheads[i],
                rdepth[i], i != nb - 1))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, boc[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(boc[0], cfg.out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None,
                return_dict=True, **unused):
        cfg = self.config
        b = sample.shape[0]
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], device=sample.device)
        t = timestep.reshape(-1).to(sample.device)
        t = t.expand(b) if t.numel() == 1 else t
        t_emb = timestep_embedding(t, cfg.block_out_channels[0], cfg.flip_sin_to_cos,
                                   cfg.freq_shift).to(sample.dtype)
        emb = self.time_embedding(t_emb)
        if cfg.addition_embed_type == "text_time":
            text_embeds = added_cond_kwargs["text_embeds"]
            time_ids = added_cond_kwargs["time_ids"]
            tid = timestep_embedding(time_ids.flatten(), cfg.addition_time_embed_dim,
                                     cfg.flip_sin_to_cos, cfg.freq_shift)
            tid = tid.reshape(b, -1).to(sample.dtype)
            emb = emb + self.add_embedding(torch.cat([text_embeds, tid], dim=-1))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states)
            skips.extend(outs)
        if down_block_additional_residuals is not None:
            # ControlNet (diffusers UNet2DConditionModel.forward): every skip tensor gets its
            # residual added AFTER the down path has consumed the un-modified one
            skips = [s + r for s, r in zip(skips, down_block_additional_residuals)]
        x = self.mid_block(x, emb, encoder_hidden_states)
        if mid_block_additional_residual is not None:
            x = x + mid_block_additional_residual
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        if not return_dict:
            return (x,)
        return UNetOutput(sample=x)


def build_unet(cfg: UNetConfig, seed: int = 0, dtype=torch.float32, device="cpu",
               weight_gain: float = 1.0, randomize_affine: bool = True):
    """Seeded default-PyTorch-init UNet (no checkpoints exist offline).

    `randomize_affine` (default): every GroupNorm / LayerNorm gets gamma ~ 1 + 0.3 N(0, 1) and
    beta ~ 0.3 N(0, 1) instead of PyTorch's (1, 0), so a whole-UNet parity test sees a norm that
    is wired to the wrong consumer, a dropped beta or a mis-permuted folded LayerNorm."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = UNet2DConditionModel(cfg)
    if randomize_affine:
        with torch.no_grad():
            for mod in m.modules():
                if isinstance(mod, (nn.GroupNorm, nn.LayerNorm)):
                    mod.weight.copy_(1.0 + 0.3 * torch.randn_like(mod.weight))
                    mod.bias.copy_(0.3 * torch.randn_like(mod.bias))
    torch.random.set_rng_state(g)
    if weight_gain != 1.0:
        with torch.no_grad():
            for p in m.parameters():
                p.mul_(weight_gain)
    return m.to(device=device, dtype=dtype).eval()
