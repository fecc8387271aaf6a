"""CPU tests of the host side: drop-in surface, architecture walk, schedule builder (dry run),
weight packing, C-ABI symbol table.  No compute kernels are called (no GPU here)."""
import ctypes
import dataclasses
import os
import re

import pytest
import torch

from oracle import unet_oracle as uo
from sfast_b200 import _lib, ops
from sfast_b200.plan import PackedWeights, UNetPlan
from sfast_b200.unet_spec import param_shapes, random_state_dict, spec_from_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_compilation_config_has_reference_fields():
    from sfast.compilers.diffusion_pipeline_compiler import CompilationConfig, compile, compile_unet, compile_vae  # noqa: F401
    names = [f.name for f in dataclasses.fields(CompilationConfig.Default)]
    # /root/reference/src/sfast/compilers/diffusion_pipeline_compiler.py:65-78
    assert names == ["memory_format", "enable_jit", "enable_jit_freeze", "preserve_parameters",
                     "enable_cnn_optimization", "enable_fused_linear_geglu", "prefer_lowp_gemm",
                     "enable_xformers", "enable_cuda_graph", "enable_triton", "trace_scheduler"]
    c = CompilationConfig.Default()
    assert c.enable_jit and c.enable_jit_freeze and c.preserve_parameters and c.prefer_lowp_gemm
    assert not c.enable_xformers and not c.enable_cuda_graph and not c.enable_triton
    from sfast.compilers import stable_diffusion_pipeline_compiler as alias
    assert alias.compile is compile


def test_compile_unet_refuses_cpu_modules_loudly():
    from sfast.compilers.diffusion_pipeline_compiler import CompilationConfig, compile_unet
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    m = uo.build_unet(uo.tiny_config())
    with pytest.raises(RuntimeError, match="no CPU"):
        compile_unet(m, CompilationConfig.Default())


@pytest.mark.parametrize("cfgf", [uo.sd15_config, uo.sdxl_config, uo.tiny_config])
def test_spec_parameter_names_match_oracle_state_dict(cfgf):
    cfg = cfgf()
    shapes = param_shapes(spec_from_config(cfg))
    with torch.device("meta"):
        sd = {k: tuple(v.shape) for k, v in uo.UNet2DConditionModel(cfg).state_dict().items()}
    assert set(sd) == set(shapes)
    assert all(tuple(shapes[k]) == sd[k] for k in sd)


def test_unsupported_architecture_is_rejected():
    cfg = uo.sd15_config()
    cfg.down_block_types = ("AttnDownBlock2D",) + cfg.down_block_types[1:]
    with pytest.raises(NotImplementedError):
        spec_from_config(cfg)


def _dry_plan(cfg, batch, h, w):
    spec = spec_from_config(cfg)
    sd = {k: torch.empty(s, dtype=torch.float16, device="meta") for k, s in param_shapes(spec).items()}
    pw = PackedWeights(spec, sd, torch.float16, "meta", dry=True)
    return UNetPlan(pw, batch, h, w)


def test_sd15_plan_matches_survey_kernel_counts_and_flops():
    plan = _dry_plan(uo.sd15_config(), 2, 64, 64)
    names = [op.fn.name for op in plan.all_ops() if op.fn is not None]
    # forked stream: time-embedding chain (4 launches) + the 16 cross-attention K/V projections
    assert len(plan.side_ops) == 20
    # 61 GroupNorms, 48 LayerNorms, 32 attention calls, 3 upsamples
    # B = 2 tensors fit in shared memory: every GroupNorm that is its own kernel takes the single-launch
    # fused one; the others are FOLDED INTO the 3x3 conv that consumes them (statistics -> (scale, shift),
    # applied on the conv's operand path): at B = 2 the three wide (960 / 640-channel) full-launch convs of the last up block
    n_folded = names.count("sfb_group_norm_scale_shift")
    assert names.count("sfb_group_norm_fused") + n_folded == 61 and names.count("sfb_group_norm_apply") == 0
    assert n_folded == (0 if ops.CONV_GN == "0" else 3)
    # all 48 LayerNorms are folded into the consuming GEMMs (gamma-scaled weights + epilogue)
    assert names.count("sfb_layer_norm") == 0
    gemms = [op.keep[0] for op in plan.all_ops() if op.fn is not None and op.fn.name == "sfb_gemm"]
    assert sum(1 for g in gemms if g.a_mode == _lib.A_CONV3X3_GN) == n_folded
    for g in gemms:
        if g.a_mode == _lib.A_CONV3X3_GN:  # what sfb_gemm demands of the halo conv
            assert (g.box_n, g.box_h, g.box_w, g.splits, g.cta_pair, g.persistent) == (1, 16, 8, 1, 1, 0)
            assert g.img_w % 8 == 0 and g.cin % 64 == 0 and g.conv_stride == 1 and g.epi == _lib.EPI_STORE
    assert sum(1 for g in gemms if g.ln_rowstats is not None or g.ln_dim > 0) == 48
    assert names.count("sfb_attention") == 32
    # nearest-2x upsample + conv3x3 runs as four 2x2 convolutions on the low-res image: the
    # upsampled tensor is never materialised
    assert names.count("sfb_upsample2x") == 0
    assert sum(1 for g in gemms if g.a_mode == _lib.A_UPCONV2X) == 3
    # 98 convs + 184 GEMMs of the reference collapse to 209 GEMM launches (fused QKV / KV, 22
    # time projections in one small_linear, conv_in as im2col + GEMM, conv_out as an edge kernel)
    assert names.count("sfb_gemm") == 209 and names.count("sfb_im2col_in") == 1
    # algorithmic FLOPs (conv/linear/attention MACs * 2), SURVEY.md section 8d: 1.607 TFLOP at B = 2
    assert abs(plan.flops() / 1e12 - 1.607) < 0.003
    # every skip concat is zero-copy: 12 concat buffers, no copy op exists
    assert sum(1 for k in plan._bufs if k[0].startswith("cat_")) == 12


@pytest.mark.parametrize("cfg_name,batch,size", [("sd15_config", 2, 64), ("sd15_config", 8, 64),
                                                 ("sd15_config", 8, 128), ("sdxl_config", 8, 128)])
def test_folded_group_norms_feed_exactly_the_conv_that_applies_them(cfg_name, batch, size):
    """Every GroupNorm folded into a conv is a (statistics -> scale/shift) launch whose buffer is consumed by
    the very next SFB_A_CONV3X3_GN conv of the main stream, over the same raw tensor; buffers are not shared
    between GroupNorm sites; the conv is never a split-K launch (its producer's partials would have nobody to
    finish them) and nothing is left pending in front of it."""
    from sfast_b200.plan import _ForkOp, _JoinOp
    plan = _dry_plan(getattr(uo, cfg_name)(), batch, size, size)
    main = [op for op in plan.ops if not isinstance(op, (_ForkOp, _JoinOp))]
    seen, n = set(), 0
    for i, op in enumerate(main):
        if op.fn is None or op.fn.name != "sfb_group_norm_scale_shift":
            continue
        n += 1
        gp, ab = op.keep[0], op.keep[-1]
        assert id(ab) not in seen and tuple(ab.shape) == (gp.n, gp.c, 2)
        seen.add(id(ab))
        nxt = next(o for o in main[i + 1:] if o.fn is not None and o.fn.name == "sfb_gemm")
        g = nxt.keep[0]
        assert g.a_mode == _lib.A_CONV3X3_GN and any(k is ab for k in nxt.keep[-3]), (op.name, nxt.name)
        assert (g.img_n, g.img_h * g.img_w, g.cin) == (gp.n, gp.hw, gp.c) and g.gn_silu == 1
        assert g.splits == 1 and g.defer_finish == 0 and g.cta_pair == 1
        assert (g.img_n * -(-g.img_h // 16) * (g.img_w // 8)) % 2 == 0
    n_conv = sum(1 for o in main if o.fn is not None and o.fn.name == "sfb_gemm"
                 and o.keep[0].a_mode == _lib.A_CONV3X3_GN)
    assert n == n_conv and (n > 0 or ops.CONV_GN == "0")
    # the kernels that were replaced are gone, not duplicated: one GroupNorm launch set per site
    spec_gn = sum(2 for _ in plan.spec.all_resnets()) + 1 + sum(
        1 for blk in plan.spec.down + [plan.spec.mid] + plan.spec.up for t in blk.attentions if t is not None)
    names = [o.fn.name for o in plan.all_ops() if o.fn is not None]
    assert n + names.count("sfb_group_norm_fused") + names.count("sfb_group_norm_stats") == spec_gn


def test_halo_conv_shared_memory_index_algebra():
    """The address algebra of the folded-GroupNorm conv (gemm_tc.cu, HALO branch), emulated: TMA's
    128-byte swizzle of the raw [18 x 10]-pixel halo tile, the transform threads' read addresses, their
    writes into the three column-shifted copies, and the UMMA K-major SWIZZLE_128B read of tap (dy, dx) as
    "copy dx advanced by dy 1024-byte atoms" must deliver halo pixel (y + dy, x + dx) as row m = 8 y + x of
    the A tile, for every tap, row and 16-byte channel chunk."""
    HR, HC, ABUF = 18, 10, 18 * 1024

    def swz(addr):  # chunk bits [4:7) ^= row bits [7:10): what TMA writes and the tensor core reads
        return addr ^ (((addr >> 7) & 7) << 4)

    raw = {swz((hy * HC + hx) * 128 + c * 16): (hy, hx, c) for hy in range(HR) for hx in range(HC) for c in range(8)}
    copies = {}
    for et in range(256):           # the 8 epilogue warps
        c, p0 = et & 7, et >> 3
        for i in range(6):
            p = p0 + 32 * i
            if p >= HR * HC:
                continue
            hy, hx = divmod(p, HC)
            assert raw[p * 128 + ((c ^ (p & 7)) << 4)] == (hy, hx, c)        # roff[i]
            for dx in range(3):
                x = hx - dx
                if 0 <= x < 8:
                    addr = dx * ABUF + hy * 1024 + x * 128 + ((c ^ x) << 4)  # the STS address
                    assert addr not in copies
                    copies[addr] = (hy, hx, c)
    assert len(copies) == 3 * HR * 8 * 8                                     # every slot written exactly once
    for dy in range(3):
        for dx in range(3):
            start = dx * ABUF + dy * 1024                                    # descriptor start, SBO = 1024
            assert start % 1024 == 0
            for m in range(128):
                y, x = divmod(m, 8)
                for k in range(8):
                    got = copies[swz(start + (m // 8) * 1024 + (m % 8) * 128 + k * 16)]
                    assert got == (y + dy, x + dx, k), (dy, dx, m, k, got)


def _check_deferred_finishes(plan):
    from sfast_b200.plan import _ForkOp, _JoinOp
    main = [op for op in plan.ops if not isinstance(op, (_ForkOp, _JoinOp))]
    n_defer = 0
    for i, op in enumerate(main):
        if op.fn is not None and op.fn.name == "sfb_gemm" and op.keep[0].defer_finish:
            g, nxt = op.keep[0], main[i + 1]
            assert nxt.fn.name == "sfb_group_norm_fused", (op.name, nxt.name)
            gn = nxt.keep[0]
            assert gn.part_splits == g.splits > 1 and gn.part_c == g.N and gn.part_ld == g.N
            assert gn.x == g.out and gn.n * gn.hw == g.M and gn.c >= g.N
            n_defer += 1
    gn_part = sum(1 for op in main if op.fn is not None and op.fn.name == "sfb_group_norm_fused"
                  and op.keep[0].part_splits > 1)
    assert n_defer == gn_part
    return n_defer


def test_deferred_split_k_finishes_are_absorbed_by_the_next_group_norm():
    """Every GEMM launched with defer_finish must be IMMEDIATELY followed (main stream order) by a
    fused GroupNorm that reads its partials with the same split count and channel count; forked
    branches are joined before the op that consumes them."""
    from sfast_b200.plan import _ForkOp, _JoinOp
    plan = _dry_plan(uo.sd15_config(), 2, 64, 64)
    assert _check_deferred_finishes(plan) == 35
    # forks: 14 resnet shortcut GEMMs, each joined exactly once, later in the list, never split
    forks = [i for i, op in enumerate(plan.ops) if isinstance(op, _ForkOp)]
    assert len(forks) == 14
    for i in forks:
        joins = [j for j, op in enumerate(plan.ops) if isinstance(op, _JoinOp) and op.kind is plan.ops[i]]
        assert len(joins) == 1 and joins[0] > i
        assert all(o.keep[0].splits == 1 for o in plan.ops[i].branch)
    # other shapes keep the invariant (large batch: only the small low-resolution tensors still take
    # the fused GroupNorm; SDXL: deeper transformers, no attention at the top level)
    _check_deferred_finishes(_dry_plan(uo.sd15_config(), 16, 64, 64))
    _check_deferred_finishes(_dry_plan(uo.sdxl_config(), 2, 128, 128))
    _check_deferred_finishes(_dry_plan(uo.sd15_config(), 1, 128, 128))


def test_plan_scales_linearly_in_batch_and_handles_128_latents():
    f1 = _dry_plan(uo.sd15_config(), 1, 64, 64).flops()
    f8 = _dry_plan(uo.sd15_config(), 8, 64, 64).flops()
    assert abs(f8 / f1 - 8) < 1e-6
    big = _dry_plan(uo.sd15_config(), 1, 128, 128).flops()
    assert abs(big / 1e12 - 4.674) < 0.01


def test_sdxl_plan_builds():
    plan = _dry_plan(uo.sdxl_config(), 2, 128, 128)
    assert abs(plan.flops() / 2 / 1e12 - 6.761) < 0.02


def test_split_k_policy():
    # cost model: ~4.5 us per launch + 0.34 us per K block + a ~10 us reduction pass
    assert ops.choose_splits(64, 2, 45) == 1       # 64^2 resnet conv: 128 tiles, no split
    assert ops.choose_splits(16, 4, 10) == 1       # 32^2 linear, K = 640: too short to split
    assert ops.choose_splits(16, 4, 90) == 2       # 32^2 conv: 64 tiles -> 2 splits
    assert ops.choose_splits(4, 8, 20) == 1        # 16^2 linear K = 1280: a reduction costs more
    assert ops.choose_splits(4, 8, 180) == 4       # 16^2 conv K = 11520: 4-way
    s = ops.choose_splits(1, 8, 180)               # 8^2 conv: weight-bandwidth-bound, fill the SMs
    assert 8 <= s <= 18 and 8 * s <= ops.num_sms()
    assert ops.choose_splits(1, 1, 5) == 1


def test_conv_tile_box():
    assert ops.conv_tile_box(64, 64) == (1, 2, 64)
    assert ops.conv_tile_box(8, 8) == (2, 8, 8)
    assert ops.conv_tile_box(4, 4) == (8, 4, 4)
    assert ops.conv_tile_box(72, 128) == (1, 1, 128)
    # widths that do not divide 128: 2-D patches
    assert ops.conv_tile_box(96, 96) == (1, 4, 32)     # 768 x 768 images
    assert ops.conv_tile_box(152, 104) == (1, 16, 8)   # 832 x 1216 (SDXL bucket)
    assert ops.conv_tile_box(24, 24) == (1, 16, 8)
    assert ops.conv_tile_box(3, 12) == (1, 32, 4)
    assert ops.conv_tile_box(5, 25) == (1, 128, 1)


def test_plans_build_for_resolutions_that_do_not_divide_128():
    for h, w in ((96, 96), (64, 96), (104, 152)):
        plan = _dry_plan(uo.sd15_config(), 1, h, w)
        assert plan.flops() > 0
    _dry_plan(uo.sdxl_config(), 2, 104, 152)


def test_geglu_packing_round_trip():
    torch.manual_seed(0)
    inner, k = 200, 64
    w, b = torch.randn(2 * inner, k), torch.randn(2 * inner)
    wp, bp, n = ops.pack_geglu(w, b, torch.float32)
    assert n == inner and wp.shape == (3 * 160, k)
    x = torch.randn(5, k)
    y = x @ wp.t() + bp
    out = torch.zeros(5, inner)
    for t in range(3):
        lo, hi = t * 80, min((t + 1) * 80, inner)
        v, g = y[:, t * 160:t * 160 + hi - lo], y[:, t * 160 + 80:t * 160 + 80 + hi - lo]
        out[:, lo:hi] = v * torch.nn.functional.gelu(g)
    h, g = (x @ w.t() + b).chunk(2, -1)
    torch.testing.assert_close(out, h * torch.nn.functional.gelu(g), rtol=1e-4, atol=1e-4)
    assert float(wp[inner % 80 + 160 * 2:160 * 2 + 80].abs().max()) == 0  # padding rows are zero


def test_conv_weight_packing_order():
    w = torch.arange(2 * 3 * 9, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = ops.pack_conv3x3(w, torch.float32)
    assert p.shape == (2, 27)
    # K index = (kh * 3 + kw) * cin + c
    assert p[1, (1 * 3 + 2) * 3 + 1] == w[1, 1, 1, 2]


def test_upconv_weight_packing_equals_upsample_then_conv():
    """ops.pack_upconv: the 4-phase 2x2 formulation reproduces conv3x3(nearest_upsample_2x(x))."""
    torch.manual_seed(0)
    cout, cin, h, w = 5, 3, 6, 7
    wt, x = torch.randn(cout, cin, 3, 3), torch.randn(1, cin, h, w)
    ref = torch.nn.functional.conv2d(torch.nn.functional.interpolate(x, scale_factor=2, mode="nearest"),
                                     wt, padding=1)
    packed = ops.pack_upconv(wt, torch.float32)
    npad = packed.shape[0] // 4
    assert packed.shape == (4 * 160, 4 * cin) and npad == 160
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))  # zero padding = TMA out-of-bounds fill
    out = torch.zeros(1, cout, 2 * h, 2 * w)
    for py in range(2):
        for px in range(2):
            wp = packed[(2 * py + px) * npad:(2 * py + px) * npad + cout].reshape(cout, 2, 2, cin)
            for ty in range(2):
                for tx in range(2):
                    dy, dx = py - 1 + ty, px - 1 + tx
                    src = xp[:, :, 1 + dy:1 + dy + h, 1 + dx:1 + dx + w]       # x[y + dy, x + dx]
                    out[:, :, py::2, px::2] += torch.einsum("oc,bchw->bohw", wp[:, ty, tx], src)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


def test_c_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sfb200.h")).read()
    declared = set(re.findall(r"\b(sfb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    h = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(h, name), name
    lib = _lib.lib()
    assert lib.sfb_abi_version() == _lib.ABI_VERSION == 4
    # argument validation works without a GPU and never falls back silently
    p = _lib.GemmParams()
    assert lib.sfb_gemm(ctypes.byref(p), None) < 0
    assert b"sfb_gemm" in lib.sfb_last_error()


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every parameter struct as gcc lays out include/sfb200.h must equal the
    ctypes mirror in _lib.py (field for field)."""
    import subprocess
    structs = {"sfb_gemm_params": _lib.GemmParams, "sfb_attn_params": _lib.AttnParams,
               "sfb_gn_params": _lib.GnParams, "sfb_ln_params": _lib.LnParams,
               "sfb_small_linear_params": _lib.SmallLinearParams}
    lines = []
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "sfb200.h"\nint main(void) {\n'
                   + "\n".join(lines) + "\nreturn 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = dict(line.split() for line in out.strip().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_random_state_dict_is_seeded_and_complete():
    spec = spec_from_config(uo.tiny_config())
    a = random_state_dict(spec, seed=3)
    b = random_state_dict(spec, seed=3)
    assert set(a) == set(param_shapes(spec))
    assert all(torch.equal(a[k], b[k]) for k in a)
    m = uo.build_unet(uo.tiny_config())
    m.load_state_dict({k: v.float() for k, v in a.items()})


# ---------------------------------------------------------------------------------------------
# Host-side validation of the C ABI on CPU: every sfb_gemm / sfb_group_norm_fused / sfb_attention
# parameter block of a (dry) plan is handed to the real library with dummy non-null pointers.  On
# a box without a GPU the call must get PAST argument validation and fail at the CUDA launch
# (SFB_ERR_CUDA), never with SFB_ERR_INVALID -- geometry the kernels refuse is caught here.
# ---------------------------------------------------------------------------------------------
_DUMMY = ctypes.create_string_buffer(256)
_DUMMY_PTR = (ctypes.addressof(_DUMMY) + 63) & ~63


def _clone(struct):
    c = type(struct)()
    ctypes.memmove(ctypes.byref(c), ctypes.byref(struct), ctypes.sizeof(struct))
    return c


def _nonnull(struct, fields):
    for f in fields:
        if not getattr(struct, f):
            setattr(struct, f, _DUMMY_PTR)


def _validate_plan_on_cpu(plan):
    lib = _lib.lib()
    n = collections_counter()
    for op in plan.all_ops():
        if op.fn is None:
            continue
        name = op.fn.name
        if name == "sfb_gemm":
            p = _clone(op.keep[0])
            _nonnull(p, ["tmap_a", "tmap_b", "out"])
            if p.splits > 1:
                _nonnull(p, ["ws"])
            if p.a_mode == _lib.A_CONV3X3_GN:
                _nonnull(p, ["gn_scale_shift"])
            rc = lib.sfb_gemm(ctypes.byref(p), None)
        elif name == "sfb_group_norm_scale_shift":
            p = _clone(op.keep[0])
            _nonnull(p, ["x", "gamma", "beta", "stats", "sync_counter"])
            rc = lib.sfb_group_norm_scale_shift(ctypes.byref(p), _DUMMY_PTR, None)
        elif name == "sfb_group_norm_fused":
            p = _clone(op.keep[0])
            _nonnull(p, ["x", "y", "gamma", "beta", "stats", "sync_counter"])
            if p.part_splits > 1:
                _nonnull(p, ["part_ws"])
            rc = lib.sfb_group_norm_fused(ctypes.byref(p), None)
        elif name in ("sfb_group_norm_stats", "sfb_group_norm_apply"):
            p = _clone(op.keep[0])
            _nonnull(p, ["x", "y", "gamma", "beta", "stats"])
            rc = getattr(lib, name)(ctypes.byref(p), None)
        elif name == "sfb_attention":
            p = _clone(op.keep[0])
            _nonnull(p, ["tmap_q", "tmap_k", "tmap_vt", "out"])
            rc = lib.sfb_attention(ctypes.byref(p), None)
        elif name == "sfb_temporal_attention":
            p = _clone(op.keep[0])
            _nonnull(p, ["qkv", "out"])
            rc = lib.sfb_temporal_attention(ctypes.byref(p), None)
        elif name in ("sfb_row_broadcast_add", "sfb_alpha_blend"):
            p = _clone(op.keep[0])
            _nonnull(p, ["x", "y", "vec" if name == "sfb_row_broadcast_add" else "x2"])
            if name == "sfb_alpha_blend":
                _nonnull(p, ["mix_factor"])
            rc = getattr(lib, name)(ctypes.byref(p), None)
        else:
            continue
        assert rc != 0, f"{op.name}: launched without a GPU?"
        assert rc == -2, f"{op.name} ({name}): host validation refused it: {lib.sfb_last_error().decode()}"
        n[name] += 1
    return n


def collections_counter():
    import collections
    return collections.Counter()


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU: launches must fail")
def test_host_validation_really_rejects_bad_geometry():
    lib = _lib.lib()
    plan = _dry_plan(uo.tiny_config(), 2, 32, 32)
    g = next(op.keep[0] for op in plan.all_ops() if op.fn is not None and op.fn.name == "sfb_gemm"
             and op.keep[0].a_mode == _lib.A_CONV3X3)
    for field, bad in (("K", 63), ("box_h", 3), ("N", 13), ("splits", 10 ** 6)):
        p = _clone(g)
        _nonnull(p, ["tmap_a", "tmap_b", "out", "ws"])
        setattr(p, field, bad)
        assert lib.sfb_gemm(ctypes.byref(p), None) == -1, field


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU: launches must fail")
@pytest.mark.parametrize("cfgf,batch,h,w", [
    (uo.sd15_config, 2, 64, 64), (uo.sd15_config, 1, 64, 64), (uo.sd15_config, 3, 64, 64),
    (uo.sd15_config, 16, 64, 64), (uo.sd15_config, 1, 128, 128), (uo.sd15_config, 2, 96, 96),
    (uo.sd15_config, 1, 64, 96), (uo.sdxl_config, 2, 128, 128), (uo.sdxl_config, 1, 104, 152),
    (uo.tiny_config, 2, 32, 32), (uo.tiny_config, 1, 24, 40),
    (uo.sd15_config, 8, 64, 64), (uo.sd15_config, 8, 128, 128), (uo.sdxl_config, 8, 128, 128),  # the bench shapes
])
def test_every_launch_of_a_plan_passes_host_validation(cfgf, batch, h, w):
    plan = _dry_plan(cfgf(), batch, h, w)
    n = _validate_plan_on_cpu(plan)
    assert n["sfb_gemm"] > 50 and n["sfb_attention"] > 0


@pytest.mark.parametrize("key,value", [("act_fn", "gelu"), ("conv_in_kernel", 5), ("num_class_embeds", 10),
                                       ("resnet_out_scale_factor", 2.0), ("attention_bias", True),
                                       ("cross_attention_norm", "layer_norm"), ("time_embedding_type", "fourier")])
def test_unsupported_config_values_are_rejected_by_name(key, value):
    from sfast_b200.synthetic import SD15
    from sfast_b200.unet_spec import spec_from_config
    with pytest.raises(NotImplementedError, match=key):
        spec_from_config(dict(SD15, **{key: value}))


def test_packed_weights_refresh_keeps_storage_and_tracks_new_values():
    """LoRA contract (reference preserve_parameters=True): packed copies are refreshed IN PLACE."""
    import torch
    from sfast_b200.plan import PackedWeights
    from sfast_b200.synthetic import TINY
    from sfast_b200.unet_spec import random_state_dict, spec_from_config
    spec = spec_from_config(TINY)
    sd = random_state_dict(spec, seed=3, dtype=torch.float32)
    pw = PackedWeights(spec, sd, torch.float16, "cpu", dry=True)  # dry: no TMA maps (no driver here)
    name = "down_blocks.0.resnets.0.conv1.weight"
    m = pw.conv3x3(name)
    lm = pw.ln_matrix(["down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.weight"],
                      "down_blocks.0.attentions.0.transformer_blocks.0.norm1")
    ptrs = (m.data.data_ptr(), lm[0].data.data_ptr(), lm[1].data_ptr(), pw.tproj_w.data_ptr())
    before = m.data.clone()
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2[name] = sd2[name] * 2.0
    pw.refresh(sd2)
    assert (m.data.data_ptr(), lm[0].data.data_ptr(), lm[1].data_ptr(), pw.tproj_w.data_ptr()) == ptrs
    torch.testing.assert_close(m.data.float(), before.float() * 2.0, rtol=2e-3, atol=1e-6)


def test_controlnet_plan_adds_one_launch_and_thirteen_static_inputs():
    import torch
    from sfast_b200.plan import PackedWeights, UNetPlan
    from sfast_b200.synthetic import SD15
    from sfast_b200.unet_spec import param_shapes, spec_from_config
    spec = spec_from_config(SD15)
    sd = {k: torch.empty(v, device="meta") for k, v in param_shapes(spec).items()}
    w = PackedWeights(spec, sd, torch.float16, "meta", dry=True)
    base = UNetPlan(w, 2, 64, 64)
    ctl = UNetPlan(w, 2, 64, 64, controlnet=True)
    assert len(ctl.all_ops()) == len(base.all_ops()) + 1
    assert len(ctl.ctrl_in) == 13 and not base.ctrl_in
    shapes = [tuple(t.shape) for t in ctl.ctrl_in]
    assert shapes[0] == (2, 320, 64, 64) and shapes[-2] == (2, 1280, 8, 8) and shapes[-1] == (2, 1280, 8, 8)


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU: launches must fail")
@pytest.mark.parametrize("tiny,videos,h,w", [(True, 2, 16, 16), (True, 3, 24, 40), (False, 2, 72, 128), (False, 1, 40, 64)])
def test_every_launch_of_an_svd_plan_passes_host_validation(tiny, videos, h, w):
    from oracle import svd_oracle as so
    from sfast_b200.svd_plan import SVDPlan
    from sfast_b200.plan import PackedWeights
    from sfast_b200.unet_spec import param_shapes, spec_from_config
    spec = spec_from_config(so.svd_tiny_config() if tiny else so.svd_xt_config())
    sd = {k: torch.empty(v, dtype=torch.float16, device="meta") for k, v in param_shapes(spec).items()}
    plan = SVDPlan(PackedWeights(spec, sd, torch.float16, "meta", dry=True), videos, h, w)
    n = _validate_plan_on_cpu(plan)
    assert n["sfb_gemm"] > 100 and n["sfb_temporal_attention"] == 16 and n["sfb_row_broadcast_add"] == 48


def test_svd_plan_builds_and_counts_its_temporal_ops():
    import torch
    from oracle import svd_oracle as so
    from sfast_b200.svd_plan import SVDPlan
    from sfast_b200.plan import PackedWeights
    from sfast_b200.unet_spec import param_shapes, spec_from_config
    cfg = so.svd_xt_config()
    spec = spec_from_config(cfg)
    assert spec.temporal and spec.num_frames == 25
    shapes = param_shapes(spec)
    assert sum(int(torch.Size(s).numel()) for s in shapes.values()) == 1_524_623_082
    sd = {k: torch.empty(v, dtype=torch.float16, device="meta") for k, v in shapes.items()}
    pw = PackedWeights(spec, sd, torch.float16, "meta", dry=True)
    plan = SVDPlan(pw, 2, 72, 128)
    names = [op.fn.name for op in plan.all_ops() if op.fn is not None]
    # 16 spatio-temporal transformers: one temporal attention, 3 broadcast adds (spatial ctx, frame
    # position, temporal ctx) and one AlphaBlender each; 22 res blocks: 2 temporal convs each
    assert names.count("sfb_temporal_attention") == 16
    assert names.count("sfb_row_broadcast_add") == 48 and names.count("sfb_alpha_blend") == 16
    gemms = [op.keep[0] for op in plan.all_ops() if op.fn is not None and op.fn.name == "sfb_gemm"]
    assert sum(1 for g in gemms if g.a_mode == _lib.A_CONV3X1) == 44
    assert names.count("sfb_attention") == 16          # spatial self-attention only: cross-attention is an add
    assert plan.flops() / 1e12 > 50                    # ~10^2 TFLOP per denoising step of a 2 x 25-frame clip
    print("SVD-XT 2x25x72x128: %.1f TFLOP / step, %d launches" % (plan.flops() / 1e12, len(names)))


def test_vae_decoder_plan_builds_and_passes_host_validation():
    import torch
    from oracle import vae_oracle as vo
    from sfast_b200.plan import PackedWeights
    from sfast_b200.vae_plan import VAEDecodePlan, vae_decoder_param_shapes, vae_spec_from_config
    spec = vae_spec_from_config(vo.sd_vae_config())
    shapes = vae_decoder_param_shapes(spec)
    assert sum(int(torch.Size(s).numel()) for s in shapes.values()) == 49_490_199  # decoder + post_quant_conv
    with torch.device("meta"):
        m = vo.AutoencoderKLDecoder(vo.sd_vae_config())
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v) for k, v in shapes.items()}
    sd = {k: torch.empty(v, dtype=torch.float16, device="meta") for k, v in shapes.items()}
    plan = VAEDecodePlan(PackedWeights(spec, sd, torch.float16, "meta", dry=True), 1, 64, 64)
    names = [op.fn.name for op in plan.all_ops() if op.fn is not None]
    assert names.count("sfb_row_softmax") == 1 and names.count("sfb_pointwise_nchw") == 1
    gemms = [op.keep[0] for op in plan.all_ops() if op.fn is not None and op.fn.name == "sfb_gemm"]
    assert sum(1 for g in gemms if g.b_plain) == 3 and sum(1 for g in gemms if g.a_mode == _lib.A_UPCONV2X) == 3
    print("VAE decode 64x64 -> 512x512: %.2f TFLOP, %d launches" % (plan.flops() / 1e12, len(names)))
    if not torch.cuda.is_available():
        n = _validate_plan_on_cpu(plan)
        assert n["sfb_gemm"] >= 35


# ---------------------------------------------------------------------------------------------
# CLIP text encoder schedule (row f4): dry plans on CPU
# ---------------------------------------------------------------------------------------------
def test_clip_text_plan_dry_launch_list_and_parameter_totals():
    from collections import Counter
    from sfast_b200 import clip_plan as cp
    from sfast_b200.plan import PackedWeights
    cfg = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
               num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu", eos_token_id=2,
               projection_dim=768, layer_norm_eps=1e-5)
    spec = cp.clip_text_spec_from_config(cfg)
    shapes = cp.clip_text_param_shapes(spec)
    assert sum(torch.Size(s).numel() for s in shapes.values()) == 123_060_480      # CLIP ViT-L/14 text model
    sd = {k: torch.empty(v, device="meta", dtype=torch.float16) for k, v in shapes.items()}
    plan = cp.ClipTextPlan(PackedWeights(spec, sd, torch.float16, "meta", dry=True), 2, 77)
    names = Counter(op.name.rsplit(".", 1)[-1] for op in plan.all_ops())
    # 5 launches per layer (LayerNorms folded, activation and residuals in epilogues) + 4 edge kernels
    assert names["qkv"] == names["core"] == names["out_proj"] == names["fc1"] == names["fc2"] == 12
    assert len(plan.all_ops()) == 5 * 12 + 4
    assert len(plan.hidden) == 13
    # SDXL text_encoder_2 (OpenCLIP bigG text tower): 694,659,840 parameters with the projection
    big = cp.clip_text_spec_from_config(dict(cfg, hidden_size=1280, intermediate_size=5120, num_hidden_layers=32,
                                             num_attention_heads=20, hidden_act="gelu", projection_dim=1280),
                                        with_projection=True)
    assert sum(torch.Size(s).numel() for s in cp.clip_text_param_shapes(big).values()) == 694_659_840
    with pytest.raises(NotImplementedError):
        cp.clip_text_spec_from_config(dict(cfg, hidden_act="relu"))
    with pytest.raises(NotImplementedError):
        cp.clip_text_spec_from_config(dict(cfg, num_attention_heads=16))      # head_dim 48


def test_clip_vision_plan_dry_launch_list_and_parameter_totals():
    from sfast_b200 import clip_plan as cp
    from sfast_b200.plan import PackedWeights
    cfg = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16,
               image_size=224, patch_size=14, hidden_act="gelu", projection_dim=1024)
    spec = cp.clip_vision_spec_from_config(cfg, with_projection=True)
    shapes = cp.clip_vision_param_shapes(spec)
    assert sum(torch.Size(s).numel() for s in shapes.values()) == 632_076_800   # OpenCLIP ViT-H/14 vision + projection
    sd = {k: torch.empty(v, device="meta", dtype=torch.float16) for k, v in shapes.items()}
    plan = cp.ClipVisionPlan(PackedWeights(spec, sd, torch.float16, "meta", dry=True), 1)
    assert spec.tokens == 257 and len(plan.hidden) == 33
    # 5 launches per layer + layer 0's stand-alone LayerNorm + 8 edge kernels at batch 1
    assert len(plan.all_ops()) == 5 * 32 + 1 + 8
