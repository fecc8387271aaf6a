"""Launch schedule of the VAE decoder (diffusers ``AutoencoderKL.decode``), the module the reference's
`compile_vae` wraps (/root/reference/src/sfast/compilers/diffusion_pipeline_compiler.py:154-190).

Same kernel families as the UNet at larger spatial sizes: GroupNorm(+SiLU) -> implicit-GEMM 3x3
convolutions with fused bias / residual, nearest-2x upsample folded into four 2x2 phase convolutions,
and ONE single-head attention over the 64 x 64 latent tokens with head_dim 512.  That head does not fit
the flash kernel's TMEM budget (O would need 528 columns), and it is one layer of ~34 GFLOP, so it runs
as five tcgen05 GEMMs around a row-softmax kernel: Q (scale folded into W_q), K, V^T (= W_v X^T: the
weight is the A operand, the activation the B operand), S = Q K^T (activation x activation: plain
row-major B operand), P = softmax(S), O = P V^T^T + b_v (rows of P sum to 1, so V's bias is added after
the product), out-projection + residual.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Tuple

import torch

from . import _lib, ops
from .ops import Act, Op, _ptr
from .plan import UNetPlan, _WsToken
from .unet_spec import cfg_get


@dataclass
class VAESpec:
    latent_channels: int
    out_channels: int
    block_out_channels: Tuple[int, ...]
    layers_per_block: int
    groups: int
    eps: float = 1e-6
    num_frames: int = 0

    @property
    def temporal(self):
        return False

    def all_resnets(self):
        return []


def vae_spec_from_config(cfg) -> VAESpec:
    boc = tuple(cfg_get(cfg, "block_out_channels"))
    for key, allowed in (("act_fn", ("silu", "swish")), ("mid_block_add_attention", (True,)),
                         ("use_quant_conv", (True,)), ("use_post_quant_conv", (True,))):
        v = cfg_get(cfg, key, allowed[0])
        if v not in allowed:
            raise NotImplementedError(f"VAE config {key}={v!r} is not supported by the B200 path")
    for t in cfg_get(cfg, "up_block_types", ("UpDecoderBlock2D",) * len(boc)):
        if t != "UpDecoderBlock2D":
            raise NotImplementedError(f"VAE up block type {t}")
    if boc[-1] % 64 or any(c % 64 for c in boc):
        raise NotImplementedError("VAE channel counts must be multiples of 64")
    return VAESpec(latent_channels=cfg_get(cfg, "latent_channels", 4), out_channels=cfg_get(cfg, "out_channels", 3),
                   block_out_channels=boc, layers_per_block=cfg_get(cfg, "layers_per_block", 2),
                   groups=cfg_get(cfg, "norm_num_groups", 32))


def vae_decoder_param_shapes(spec: VAESpec):
    """{diffusers parameter name: shape} of post_quant_conv + decoder."""
    out = {}

    def conv(p, cout, cin, k):
        out[p + ".weight"] = (cout, cin, k, k)
        out[p + ".bias"] = (cout,)

    def norm(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)

    def lin(p, n, k):
        out[p + ".weight"] = (n, k)
        out[p + ".bias"] = (n,)

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)

    lc, boc = spec.latent_channels, spec.block_out_channels
    conv("post_quant_conv", lc, lc, 1)
    c = boc[-1]
    conv("decoder.conv_in", c, lc, 3)
    resnet("decoder.mid_block.resnets.0", c, c)
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        lin(f"{a}.{n}", c, c)
    resnet("decoder.mid_block.resnets.1", c, c)
    cout = c
    for i, ch in enumerate(reversed(boc)):
        cin, cout = cout, ch
        for j in range(spec.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i != len(boc) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    norm("decoder.conv_norm_out", boc[0])
    conv("decoder.conv_out", spec.out_channels, boc[0], 3)
    return out


class VAEDecodePlan(UNetPlan):
    """Static launch schedule of `AutoencoderKL.decode` for a fixed (batch, latent height, width)."""

    def __init__(self, weights, batch, height, width):  # noqa: super().__init__ is UNet-specific
        self.controlnet, self.ctrl_in = False, []
        self.w, self.spec = weights, weights.spec
        self.dt, self.dev, self.dry = weights.dtype, weights.device, weights.dry
        self.lib = None if self.dry else _lib.lib()
        self.B, self.H, self.W, self.ctx_len = batch, height, width, 0
        self.ops, self.side_ops = [], []
        self._side_stream, self._pending, self._gn_ws_patches, self._joined = None, None, [], True
        self._bufs, self._gn_count = {}, 0
        spec = self.spec
        self.z_in = self._alloc((batch, spec.latent_channels, height, width), self.dt)
        up = 1 << (len(spec.block_out_channels) - 1)
        self.out = self._alloc((batch, spec.out_channels, height * up, width * up), self.dt)
        n_gn = 2 * (2 + len(spec.block_out_channels) * (spec.layers_per_block + 1)) + 2
        self._alloc_gn(n_gn, batch)
        self.ln_arena = self._alloc((2,), torch.float32)
        self._ln_used, self._ln_slots = 0, []
        self.ws, self._ws_need = None, 0
        self._build()

    def vae_resnet(self, p, x: Act, dst: Act, cin, cout):
        eps = self.spec.eps
        a1, gn1 = self.norm_for_conv(p + ".norm1", x, p + ".norm1", eps, cout)
        h1 = self.act("res_h1", x.n, x.h, x.w, cout)
        self.conv3x3(p + ".conv1", a1, p + ".conv1", h1, gn=gn1)
        a2, gn2 = self.norm_for_conv(p + ".norm2", h1, p + ".norm2", eps, cout)
        res = x
        if cin != cout:
            res = self.act("res_sc", x.n, x.h, x.w, cout)
            self.linear(p + ".conv_shortcut", x, self.w.matrix(p + ".conv_shortcut.weight"),
                        self.w.f32(p + ".conv_shortcut.bias"), res)
        self.conv3x3(p + ".conv2", a2, p + ".conv2", dst, residual=res, gn=gn2)

    def vae_attention(self, p, x: Act, dst: Act):
        """dst = x + to_out(softmax(q k^T / sqrt(c)) v) over the h*w tokens of each image, one head."""
        lib, c, S, B = self.lib_or_dry(), x.c, x.h * x.w, x.n
        hn = self.group_norm(p + ".group_norm", x, p + ".group_norm", False, self.spec.eps)
        q = self.act("vae_q", x.n, x.h, x.w, c)
        k = self.act("vae_k", x.n, x.h, x.w, c)
        wq, bq = self.w.scaled_linear(p + ".to_q", c ** -0.5)
        self.linear(p + ".to_q", hn, wq, bq, q)
        self.linear(p + ".to_k", hn, self.w.matrix(p + ".to_k.weight"), self.w.f32(p + ".to_k.bias"), k)
        vt = self.buf("vae_vt", (B, c, S))        # V^T per image: [c, S]
        # fp32 scores, one image at a time (stream order makes reuse safe); the softmax writes the 16-bit
        # probabilities over the first half of each row (16-bit scores cost ~2 % on the probabilities)
        sc = self.buf("vae_scores", (S, S), torch.float32)
        o = self.act("vae_o", x.n, x.h, x.w, c)
        wv_plain = self.w.small(p + ".to_v.weight")  # [c, c] row-major: the A operand of V^T = W_v X^T
        es = 2
        for b in range(B):
            hn_b = _ptr(hn.buf) + b * S * hn.ld * es
            # V^T[c, S] = W_v [c, c] . hn_b[S, c]^T   (bias added after P V: softmax rows sum to 1)
            self._emit(self._gemm(f"{p}.to_v^T[{b}]", a=ops.a_matrix(_ptr(wv_plain), c, c, c),
                                  b=ops.PlainB(hn_b, S, c, hn.ld, self.dry), M=c, N=S, K=c, dt=self.dt,
                                  out=_ptr(vt) + b * c * S * es, ldo=S, splits=1, keep=(wv_plain, hn.buf, vt)))
            # S = q_b k_b^T
            self._emit(self._gemm(f"{p}.qk^T[{b}]", a=ops.a_matrix(_ptr(q.buf) + b * S * c * es, S, c, c),
                                  b=ops.PlainB(_ptr(k.buf) + b * S * c * es, S, c, c, self.dry), M=S, N=S, K=c,
                                  dt=self.dt, out=_ptr(sc), ldo=S, splits=1, epi=_lib.EPI_STORE_F32,
                                  keep=(q.buf, k.buf, sc)))
            self._emit(Op(f"{p}.softmax[{b}]", lib.sfb_row_softmax,
                          (_ptr(sc), _ptr(sc), S, S, S, 2 * S, 1, ops.dtype_code(self.dt)), (sc,), 0, 6 * S * S))
            # O = P V + b_v   (P: 16-bit rows at a pitch of 2 * S elements inside the fp32 buffer)
            self._emit(self._gemm(f"{p}.pv[{b}]", a=ops.a_matrix(_ptr(sc), S, S, 2 * S),
                                  b=ops.PlainB(_ptr(vt) + b * c * S * es, c, S, S, self.dry), M=S, N=c, K=S,
                                  dt=self.dt, out=_ptr(o.buf) + b * S * c * es, ldo=c,
                                  bias=self.w.f32(p + ".to_v.bias"), splits=1, keep=(sc, vt, o.buf)))
        self.linear(p + ".to_out", o, self.w.matrix(p + ".to_out.0.weight"), self.w.f32(p + ".to_out.0.bias"),
                    dst, residual=x)

    def _build(self):
        spec, B, H, W, lib = self.spec, self.B, self.H, self.W, self.lib_or_dry()
        self._ws_token = _WsToken()
        self._emit(Op("gn_sync.zero", lib.sfb_memset, (_ptr(self.gn_sync), 0, self.gn_sync.numel() * 4),
                      (self.gn_sync,)))
        lc, boc = spec.latent_channels, spec.block_out_channels
        c = boc[-1]
        # post_quant_conv (1x1, NCHW -> NCHW) then conv_in (NCHW -> NHWC)
        zq = self.buf("pq_out", (B, lc, H, W))
        wpq = self.w.small("post_quant_conv.weight")
        self._emit(Op("post_quant_conv", lib.sfb_pointwise_nchw,
                      (_ptr(self.z_in), _ptr(wpq), _ptr(self.w.f32("post_quant_conv.bias")), _ptr(zq), B, H * W,
                       lc, lc, ops.dtype_code(self.dt)), (self.z_in, wpq, zq)))
        x = self.act("vae_x0", B, H, W, c)
        w_in = self.w.conv_in_weight("decoder.conv_in.weight")
        self._emit(Op("decoder.conv_in", lib.sfb_conv_in,
                      (_ptr(zq), _ptr(w_in), _ptr(self.w.f32("decoder.conv_in.bias")), x.ptr, B, H, W, lc, c,
                       x.ld, ops.dtype_code(self.dt)), (zq, w_in, x.buf), 2 * B * H * W * c * 9 * lc))
        m = "decoder.mid_block"
        a = self.act("vae_x1", B, H, W, c)
        self.vae_resnet(m + ".resnets.0", x, a, c, c)
        self.vae_attention(m + ".attentions.0", a, x)
        self.vae_resnet(m + ".resnets.1", x, a, c, c)
        x, cout, h, w = a, c, H, W
        for i, ch in enumerate(reversed(boc)):
            cin, cout = cout, ch
            for j in range(spec.layers_per_block + 1):
                dst = self.act(f"vae_up{i}_{j % 2}", B, h, w, cout)
                self.vae_resnet(f"decoder.up_blocks.{i}.resnets.{j}", x, dst, cin if j == 0 else cout, cout)
                x = dst
            if i != len(boc) - 1:
                dst = self.act(f"vae_ups{i}", B, 2 * h, 2 * w, cout)
                self.upconv3x3(f"decoder.up_blocks.{i}.upsamplers.0.conv", x,
                               f"decoder.up_blocks.{i}.upsamplers.0.conv", dst)
                x, h, w = dst, 2 * h, 2 * w
        y = self.group_norm("decoder.conv_norm_out", x, "decoder.conv_norm_out", True, spec.eps)
        w_out = self.w.conv3x3_plain("decoder.conv_out.weight")
        self._emit(Op("decoder.conv_out", lib.sfb_conv_out,
                      (y.ptr, _ptr(w_out), _ptr(self.w.f32("decoder.conv_out.bias")), _ptr(self.out), B, h, w,
                       boc[0], spec.out_channels, y.ld, ops.dtype_code(self.dt)), (y.buf, w_out, self.out),
                      2 * B * h * w * boc[0] * 9 * spec.out_channels))
        assert self._gn_count <= self.gn_stats.shape[0], (self._gn_count, self.gn_stats.shape)
        if self._pending is not None:
            self._pending["p"].defer_finish = 0
            self._pending = None
        self._ws_token.finalize(self)
