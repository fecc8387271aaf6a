"""Architecture description of a diffusers ``UNet2DConditionModel`` derived from its config.

The plan builder never touches ``nn.Module`` classes: it consumes ``unet.config`` (any object or
dict with the diffusers field names) and ``unet.state_dict()`` (diffusers parameter names).  This
module turns the config into the ordered block structure both the weight packer and the plan
builder walk, and can synthesise a random state dict of that architecture for benchmarks (there
are no checkpoints offline).
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch


def cfg_get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def _tuple(v, n):
    if isinstance(v, (list, tuple)):
        assert len(v) == n
        return tuple(v)
    return (v,) * n


@dataclass
class ResnetSpec:
    prefix: str
    cin: int
    cout: int
    eps: Optional[float] = None          # GroupNorm eps (None: the UNet's norm_eps)
    # SVD SpatioTemporalResBlock: `prefix` is the spatial ResnetBlock2D; the temporal
    # TemporalResnetBlock ((3,1,1) convolutions over the frame axis) and the AlphaBlender follow
    temporal_prefix: Optional[str] = None
    mixer: Optional[str] = None          # name of the AlphaBlender's mix_factor parameter

    @property
    def has_shortcut(self):
        return self.cin != self.cout


@dataclass
class TransformerSpec:
    prefix: str
    dim: int
    heads: int
    depth: int
    ctx_dim: int
    linear_proj: bool
    temporal: bool = False  # SVD TransformerSpatioTemporalModel (temporal blocks + time_pos_embed + mixer)

    @property
    def head_dim(self):
        return self.dim // self.heads


@dataclass
class BlockSpec:
    prefix: str
    resnets: List[ResnetSpec]
    attentions: List[Optional[TransformerSpec]]
    sampler: Optional[str] = None  # prefix of the down/up-sampler conv
    cout: int = 0


@dataclass
class UNetSpec:
    in_channels: int
    out_channels: int
    block_out_channels: tuple
    temb_dim: int
    groups: int
    eps: float
    flip_sin_to_cos: bool
    freq_shift: float
    cross_attention_dim: int
    addition_embed_type: Optional[str]
    addition_time_embed_dim: Optional[int]
    add_in_dim: Optional[int]
    down: List[BlockSpec] = field(default_factory=list)
    mid: Optional[BlockSpec] = None
    up: List[BlockSpec] = field(default_factory=list)
    num_frames: int = 0     # > 0: UNetSpatioTemporalConditionModel (SVD)

    @property
    def temporal(self):
        return self.num_frames > 0

    def all_resnets(self):
        """Every block that owns a `time_emb_proj` (SVD: spatial and temporal halves)."""
        out = []
        for b in self.down + [self.mid] + self.up:
            for r in b.resnets:
                out.append(r)
                if r.temporal_prefix:
                    out.append(ResnetSpec(r.temporal_prefix, r.cout, r.cout, r.eps))
        return out


_SUPPORTED_DOWN = ("CrossAttnDownBlock2D", "DownBlock2D")
_SUPPORTED_UP = ("CrossAttnUpBlock2D", "UpBlock2D")


def spec_from_config(cfg) -> UNetSpec:
    if is_svd_config(cfg):
        return svd_spec_from_config(cfg)
    boc = tuple(cfg_get(cfg, "block_out_channels"))
    nb = len(boc)
    down_types = tuple(cfg_get(cfg, "down_block_types"))
    up_types = tuple(cfg_get(cfg, "up_block_types"))
    for t in down_types:
        if t not in _SUPPORTED_DOWN:
            raise NotImplementedError(f"down block type {t} is not supported by the B200 path")
    for t in up_types:
        if t not in _SUPPORTED_UP:
            raise NotImplementedError(f"up block type {t} is not supported by the B200 path")
    for key in ("class_embed_type", "encoder_hid_dim_type", "time_cond_proj_dim",
                "resnet_time_scale_shift", "dual_cross_attention", "only_cross_attention",
                "upcast_attention", "time_embedding_act_fn", "attention_type"):
        v = cfg_get(cfg, key)
        if v not in (None, False, "default", "positional"):
            raise NotImplementedError(f"UNet config {key}={v!r} is not supported by the B200 path")
    # fields whose only supported value is the SD-1.5 / SDXL one: anything else would compile and
    # return wrong numerics, so it is refused by name
    required = {
        "act_fn": ("silu", "swish"), "time_embedding_type": ("positional",),
        "conv_in_kernel": (3,), "conv_out_kernel": (3,), "downsample_padding": (1,),
        "mid_block_scale_factor": (1, 1.0), "resnet_out_scale_factor": (1, 1.0),
        "resnet_skip_time_act": (False,), "mid_block_only_cross_attention": (None, False),
        "cross_attention_norm": (None,), "attention_bias": (False,), "num_class_embeds": (None,),
        "addition_embed_type_num_heads": (64,), "timestep_post_act": (None,),
        "time_embedding_dim": (None,), "class_embeddings_concat": (False,),
        "reverse_transformer_layers_per_block": (None,), "dropout": (0, 0.0),
        "center_input_sample": (False,), "encoder_hid_dim": (None,),
    }
    for key, allowed in required.items():
        v = cfg_get(cfg, key, allowed[0])
        if v not in allowed:
            raise NotImplementedError(f"UNet config {key}={v!r} is not supported by the B200 path "
                                      f"(supported: {allowed})")
    mid_type = cfg_get(cfg, "mid_block_type", "UNetMidBlock2DCrossAttn")
    if mid_type not in (None, "UNetMidBlock2DCrossAttn"):
        raise NotImplementedError(f"mid block type {mid_type}")
    layers = cfg_get(cfg, "layers_per_block", 2)
    if isinstance(layers, (list, tuple)):
        if len(set(layers)) != 1:
            raise NotImplementedError("per-block layers_per_block")
        layers = layers[0]
    # `num_attention_heads` if given, else diffusers' historical alias `attention_head_dim`
    heads_cfg = cfg_get(cfg, "num_attention_heads") or cfg_get(cfg, "attention_head_dim")
    heads = _tuple(heads_cfg, nb)
    depth = _tuple(cfg_get(cfg, "transformer_layers_per_block", 1), nb)
    ctx = cfg_get(cfg, "cross_attention_dim")
    if isinstance(ctx, (list, tuple)):
        if len(set(ctx)) != 1:
            raise NotImplementedError("per-block cross_attention_dim")
        ctx = ctx[0]
    linear = bool(cfg_get(cfg, "use_linear_projection", False))
    add_type = cfg_get(cfg, "addition_embed_type")
    if add_type not in (None, "text_time"):
        raise NotImplementedError(f"addition_embed_type={add_type}")
    spec = UNetSpec(
        in_channels=cfg_get(cfg, "in_channels", 4), out_channels=cfg_get(cfg, "out_channels", 4),
        block_out_channels=boc, temb_dim=boc[0] * 4, groups=cfg_get(cfg, "norm_num_groups", 32),
        eps=cfg_get(cfg, "norm_eps", 1e-5), flip_sin_to_cos=cfg_get(cfg, "flip_sin_to_cos", True),
        freq_shift=float(cfg_get(cfg, "freq_shift", 0)), cross_attention_dim=ctx,
        addition_embed_type=add_type,
        addition_time_embed_dim=cfg_get(cfg, "addition_time_embed_dim"),
        add_in_dim=cfg_get(cfg, "projection_class_embeddings_input_dim"))

    def tf(prefix, dim, i):
        return TransformerSpec(prefix, dim, heads[i], depth[i], ctx, linear)

    cout = boc[0]
    for i, t in enumerate(down_types):
        cin, cout = cout, boc[i]
        p = f"down_blocks.{i}"
        res = [ResnetSpec(f"{p}.resnets.{j}", cin if j == 0 else cout, cout) for j in range(layers)]
        att = [tf(f"{p}.attentions.{j}", cout, i) if t.startswith("CrossAttn") else None
               for j in range(layers)]
        spec.down.append(BlockSpec(p, res, att, f"{p}.downsamplers.0.conv" if i != nb - 1 else None,
                                   cout))
    c = boc[-1]
    spec.mid = BlockSpec("mid_block",
                         [ResnetSpec("mid_block.resnets.0", c, c),
                          ResnetSpec("mid_block.resnets.1", c, c)],
                         [tf("mid_block.attentions.0", c, nb - 1)], None, c)
    rboc = list(reversed(boc))
    cout = rboc[0]
    for i, t in enumerate(up_types):
        cprev, cout = cout, rboc[i]
        cin = rboc[min(i + 1, nb - 1)]
        p = f"up_blocks.{i}"
        res = []
        for j in range(layers + 1):
            skip = cin if j == layers else cout
            rin = cprev if j == 0 else cout
            res.append(ResnetSpec(f"{p}.resnets.{j}", rin + skip, cout))
        att = [tf(f"{p}.attentions.{j}", cout, nb - 1 - i) if t.startswith("CrossAttn") else None
               for j in range(layers + 1)]
        spec.up.append(BlockSpec(p, res, att, f"{p}.upsamplers.0.conv" if i != nb - 1 else None,
                                 cout))
    return spec


_SVD_DOWN = ("CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal")
_SVD_UP = ("CrossAttnUpBlockSpatioTemporal", "UpBlockSpatioTemporal")


def is_svd_config(cfg):
    return any("SpatioTemporal" in t for t in (cfg_get(cfg, "down_block_types") or ()))


def svd_spec_from_config(cfg) -> UNetSpec:
    """diffusers UNetSpatioTemporalConditionModel (Stable Video Diffusion)."""
    boc = tuple(cfg_get(cfg, "block_out_channels"))
    nb = len(boc)
    down_types, up_types = tuple(cfg_get(cfg, "down_block_types")), tuple(cfg_get(cfg, "up_block_types"))
    for t in down_types:
        if t not in _SVD_DOWN:
            raise NotImplementedError(f"down block type {t} is not supported by the B200 path")
    for t in up_types:
        if t not in _SVD_UP:
            raise NotImplementedError(f"up block type {t} is not supported by the B200 path")
    layers = cfg_get(cfg, "layers_per_block", 2)
    if isinstance(layers, (list, tuple)):
        if len(set(layers)) != 1:
            raise NotImplementedError("per-block layers_per_block")
        layers = layers[0]
    heads = _tuple(cfg_get(cfg, "num_attention_heads"), nb)
    depth = _tuple(cfg_get(cfg, "transformer_layers_per_block", 1), nb)
    ctx = cfg_get(cfg, "cross_attention_dim")
    if isinstance(ctx, (list, tuple)):
        ctx = ctx[0]
    frames = int(cfg_get(cfg, "num_frames", 25))
    if frames > 32:
        raise NotImplementedError("temporal attention kernel handles at most 32 frames")
    for i in range(nb):
        if boc[i] % heads[i] or boc[i] // heads[i] != 64:
            raise NotImplementedError("SVD path needs attention head_dim 64")
    spec = UNetSpec(
        in_channels=cfg_get(cfg, "in_channels", 8), out_channels=cfg_get(cfg, "out_channels", 4),
        block_out_channels=boc, temb_dim=boc[0] * 4, groups=cfg_get(cfg, "norm_num_groups", 32) or 32,
        eps=1e-5, flip_sin_to_cos=True, freq_shift=0.0, cross_attention_dim=ctx,
        addition_embed_type="time_ids",
        addition_time_embed_dim=cfg_get(cfg, "addition_time_embed_dim", 256),
        add_in_dim=cfg_get(cfg, "projection_class_embeddings_input_dim", 768), num_frames=frames)

    def res(p, cin, cout, eps):
        return ResnetSpec(p + ".spatial_res_block", cin, cout, eps, p + ".temporal_res_block",
                          p + ".time_mixer.mix_factor")

    def tf(prefix, dim, i):
        return TransformerSpec(prefix, dim, heads[i], depth[i], ctx, True, temporal=True)

    cout = boc[0]
    for i, t in enumerate(down_types):
        cin, cout = cout, boc[i]
        p = f"down_blocks.{i}"
        attn = t.startswith("CrossAttn")
        eps = 1e-6 if attn else 1e-5
        rs = [res(f"{p}.resnets.{j}", cin if j == 0 else cout, cout, eps) for j in range(layers)]
        att = [tf(f"{p}.attentions.{j}", cout, i) if attn else None for j in range(layers)]
        spec.down.append(BlockSpec(p, rs, att, f"{p}.downsamplers.0.conv" if i != nb - 1 else None, cout))
    c = boc[-1]
    spec.mid = BlockSpec("mid_block", [res("mid_block.resnets.0", c, c, 1e-5), res("mid_block.resnets.1", c, c, 1e-5)],
                         [tf("mid_block.attentions.0", c, nb - 1)], None, c)
    rboc = list(reversed(boc))
    cout = rboc[0]
    for i, t in enumerate(up_types):
        cprev, cout = cout, rboc[i]
        cin = rboc[min(i + 1, nb - 1)]
        p = f"up_blocks.{i}"
        rs = []
        for j in range(layers + 1):
            skip = cin if j == layers else cout
            rin = cprev if j == 0 else cout
            rs.append(res(f"{p}.resnets.{j}", rin + skip, cout, 1e-6))
        att = [tf(f"{p}.attentions.{j}", cout, nb - 1 - i) if t.startswith("CrossAttn") else None
               for j in range(layers + 1)]
        spec.up.append(BlockSpec(p, rs, att, f"{p}.upsamplers.0.conv" if i != nb - 1 else None, cout))
    return spec


def param_shapes(spec: UNetSpec):
    """Ordered {diffusers parameter name: shape} of the architecture."""
    out = {}

    def conv(p, cout, cin, k):
        out[p + ".weight"] = (cout, cin, k, k)
        out[p + ".bias"] = (cout,)

    def lin(p, n, k, bias=True):
        out[p + ".weight"] = (n, k)
        if bias:
            out[p + ".bias"] = (n,)

    def norm(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)

    def resnet(r):
        norm(r.prefix + ".norm1", r.cin)
        conv(r.prefix + ".conv1", r.cout, r.cin, 3)
        lin(r.prefix + ".time_emb_proj", r.cout, spec.temb_dim)
        norm(r.prefix + ".norm2", r.cout)
        conv(r.prefix + ".conv2", r.cout, r.cout, 3)
        if r.has_shortcut:
            conv(r.prefix + ".conv_shortcut", r.cout, r.cin, 1)
        if r.temporal_prefix:
            t = r.temporal_prefix
            norm(t + ".norm1", r.cout)
            out[t + ".conv1.weight"] = (r.cout, r.cout, 3, 1, 1)
            out[t + ".conv1.bias"] = (r.cout,)
            lin(t + ".time_emb_proj", r.cout, spec.temb_dim)
            norm(t + ".norm2", r.cout)
            out[t + ".conv2.weight"] = (r.cout, r.cout, 3, 1, 1)
            out[t + ".conv2.bias"] = (r.cout,)
            out[r.mixer] = (1,)

    def tblock(b, dim, ctx_dim, temporal):
        if temporal:
            norm(b + ".norm_in", dim)
            lin(b + ".ff_in.net.0.proj", dim * 8, dim)
            lin(b + ".ff_in.net.2", dim, dim * 4)
        norm(b + ".norm1", dim)
        for a, kd in (("attn1", dim), ("attn2", ctx_dim)):
            lin(f"{b}.{a}.to_q", dim, dim, False)
            lin(f"{b}.{a}.to_k", dim, kd, False)
            lin(f"{b}.{a}.to_v", dim, kd, False)
            lin(f"{b}.{a}.to_out.0", dim, dim)
            if a == "attn1":
                norm(b + ".norm2", dim)
        norm(b + ".norm3", dim)
        lin(b + ".ff.net.0.proj", dim * 8, dim)
        lin(b + ".ff.net.2", dim, dim * 4)

    def transformer(t):
        if t.temporal:
            norm(t.prefix + ".norm", t.dim)
            lin(t.prefix + ".proj_in", t.dim, t.dim)
            for d in range(t.depth):
                tblock(f"{t.prefix}.transformer_blocks.{d}", t.dim, t.ctx_dim, False)
            for d in range(t.depth):
                tblock(f"{t.prefix}.temporal_transformer_blocks.{d}", t.dim, t.ctx_dim, True)
            lin(t.prefix + ".time_pos_embed.linear_1", t.dim * 4, t.dim)
            lin(t.prefix + ".time_pos_embed.linear_2", t.dim, t.dim * 4)
            out[t.prefix + ".time_mixer.mix_factor"] = (1,)
            lin(t.prefix + ".proj_out", t.dim, t.dim)
            return
        norm(t.prefix + ".norm", t.dim)
        if t.linear_proj:
            lin(t.prefix + ".proj_in", t.dim, t.dim)
        else:
            conv(t.prefix + ".proj_in", t.dim, t.dim, 1)
        for d in range(t.depth):
            b = f"{t.prefix}.transformer_blocks.{d}"
            norm(b + ".norm1", t.dim)
            for a, kd in (("attn1", t.dim), ("attn2", t.ctx_dim)):
                lin(f"{b}.{a}.to_q", t.dim, t.dim, False)
                lin(f"{b}.{a}.to_k", t.dim, kd, False)
                lin(f"{b}.{a}.to_v", t.dim, kd, False)
                lin(f"{b}.{a}.to_out.0", t.dim, t.dim)
                if a == "attn1":
                    norm(b + ".norm2", t.dim)
            norm(b + ".norm3", t.dim)
            lin(b + ".ff.net.0.proj", t.dim * 8, t.dim)
            lin(b + ".ff.net.2", t.dim, t.dim * 4)
        if t.linear_proj:
            lin(t.prefix + ".proj_out", t.dim, t.dim)
        else:
            conv(t.prefix + ".proj_out", t.dim, t.dim, 1)

    conv("conv_in", spec.block_out_channels[0], spec.in_channels, 3)
    lin("time_embedding.linear_1", spec.temb_dim, spec.block_out_channels[0])
    lin("time_embedding.linear_2", spec.temb_dim, spec.temb_dim)
    if spec.addition_embed_type in ("text_time", "time_ids"):
        lin("add_embedding.linear_1", spec.temb_dim, spec.add_in_dim)
        lin("add_embedding.linear_2", spec.temb_dim, spec.temb_dim)
    for blk in spec.down + [spec.mid] + spec.up:
        if blk is spec.mid:
            resnet(blk.resnets[0])
            transformer(blk.attentions[0])
            resnet(blk.resnets[1])
            continue
        for r, t in zip(blk.resnets, blk.attentions):
            resnet(r)
            if t is not None:
                transformer(t)
        if blk.sampler:
            conv(blk.sampler, blk.cout, blk.cout, 3)
    norm("conv_norm_out", spec.block_out_channels[0])
    conv("conv_out", spec.out_channels, spec.block_out_channels[0], 3)
    return out


def random_state_dict(spec: UNetSpec, seed=0, dtype=torch.float16, device="cpu"):
    """Random weights of the architecture (PyTorch-default-like uniform(-1/sqrt(fan_in), ..);
    norm layers: gamma ~ 1 + 0.3 N(0, 1), beta ~ 0.3 N(0, 1))."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = {}
    for name, shape in param_shapes(spec).items():
        is_norm = ".norm" in name or name.startswith("conv_norm_out")
        if name.endswith("mix_factor"):
            t = torch.randn(shape, generator=g) * 1.5
        elif is_norm:
            # random affine (not PyTorch's 1 / 0): a benchmark must not hide a dropped beta
            t = torch.randn(shape, generator=g) * 0.3 + (1.0 if name.endswith("weight") else 0.0)
        else:
            wshape = shape if name.endswith("weight") else None
            if wshape is None:
                # bias: fan-in of the matching weight
                wshape = param_shapes_cache_fan_in(spec, name)
            fan_in = int(math.prod(wshape[1:])) if len(wshape) > 1 else int(wshape[0])
            bound = 1.0 / math.sqrt(max(fan_in, 1))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = t.to(dtype=dtype, device=device)
    return sd


_FAN_CACHE = {}


def param_shapes_cache_fan_in(spec, bias_name):
    key = id(spec)
    if key not in _FAN_CACHE:
        _FAN_CACHE.clear()
        _FAN_CACHE[key] = param_shapes(spec)
    return _FAN_CACHE[key][bias_name[:-4] + "weight"]
