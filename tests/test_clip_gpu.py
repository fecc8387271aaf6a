"""GPU parity of the CLIP text encoders (`compile()` -> text_encoder / text_encoder_2 on the native path,
reference /root/reference/src/sfast/compilers/diffusion_pipeline_compiler.py:93-112).

The checker here is not a restatement: `transformers` is installed in this image, so the compiled module
is compared with THE SAME `CLIPTextModel` / `CLIPTextModelWithProjection` class a diffusers pipeline
carries -- evaluated eagerly in fp32 on the same weights (random-init; LayerNorm affines and biases
randomised so a mis-folded LayerNorm fails).  Tolerance: 1e-2 (fp16) / 4e-2 (bf16) in the elementwise +
max-norm metric of the other suites; the same module run eagerly in the 16-bit type is printed beside it.
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
transformers = pytest.importorskip("transformers")


def _rel(got, ref):
    got, ref = got.float(), ref.float()
    d = (got - ref).abs()
    rms = ref.pow(2).mean().sqrt()
    return max((d.max() / ref.abs().max()).item(), (d / (ref.abs() + rms)).max().item())


def _config(hidden, heads, layers, inter, act, eos, proj=None, vocab=49408):
    return transformers.CLIPTextConfig(
        vocab_size=vocab, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
        num_attention_heads=heads, max_position_embeddings=77, hidden_act=act, eos_token_id=eos,
        bos_token_id=vocab - 2, pad_token_id=1, projection_dim=proj or hidden)


def _build(cfg, with_projection, dtype, seed):
    torch.manual_seed(seed)
    cls = transformers.CLIPTextModelWithProjection if with_projection else transformers.CLIPTextModel
    m = cls(cfg).eval()
    with torch.no_grad():
        for name, p in m.named_parameters():
            if "layer_norm" in name:
                p.copy_(1.0 + 0.3 * torch.randn_like(p) if name.endswith("weight") else 0.3 * torch.randn_like(p))
            elif name.endswith(".bias"):
                p.copy_(0.1 * torch.randn_like(p))
    ref = copy.deepcopy(m).to("cuda", torch.float32)
    return m.to("cuda", dtype), ref


def _ids(cfg, batch, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, cfg.vocab_size - 2, (batch, 77), generator=g)
    eos = cfg.vocab_size - 1
    ids[:, 0] = cfg.vocab_size - 2
    for b in range(batch):       # SD tokenizer layout: bos, tokens, eos, then padded with eos
        ids[b, 6 + 9 * b:] = eos
    return ids.cuda()


def _compile_pipe(**mods):
    from sfast.compilers.diffusion_pipeline_compiler import CompilationConfig, compile_text_encoder
    c = CompilationConfig.Default()
    c.enable_cuda_graph = True
    return {k: compile_text_encoder(v, c) for k, v in mods.items()}


TINY = dict(hidden=128, heads=2, layers=2, inter=256)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 1e-2), (torch.bfloat16, 4e-2)])
def test_tiny_text_encoder_all_outputs(dtype, tol):
    cfg = _config(**TINY, act="quick_gelu", eos=2, vocab=1000)
    fast, ref = _build(cfg, False, dtype, seed=1)
    eager16 = copy.deepcopy(fast)
    fast = _compile_pipe(text_encoder=fast)["text_encoder"]
    ids = _ids(cfg, 3, seed=2)
    with torch.no_grad():
        got = fast(ids, output_hidden_states=True)
        want = ref(ids, output_hidden_states=True)
        lib = eager16(ids, output_hidden_states=True)
    assert type(got).__name__ == type(want).__name__
    e = _rel(got.last_hidden_state, want.last_hidden_state)
    print(f"tiny CLIP text {dtype}: native {e:.3e} | eager 16-bit {_rel(lib.last_hidden_state, want.last_hidden_state):.3e}")
    assert e < tol
    assert _rel(got.pooler_output, want.pooler_output) < tol
    assert len(got.hidden_states) == len(want.hidden_states) == cfg.num_hidden_layers + 1
    for g, w in zip(got.hidden_states, want.hidden_states):
        assert g.shape == w.shape and _rel(g, w) < tol
    assert _rel(got[0], want[0]) < tol                      # integer indexing, as pipelines do
    tup = fast(ids, return_dict=False)
    assert isinstance(tup, tuple) and len(tup) == 2 and _rel(tup[0], want.last_hidden_state) < tol


def test_replay_tracks_new_ids_fresh_outputs_and_cache():
    cfg = _config(**TINY, act="quick_gelu", eos=999, vocab=1000)
    fast, ref = _build(cfg, False, torch.float16, seed=3)
    fast = _compile_pipe(text_encoder=fast)["text_encoder"]
    a, b = _ids(cfg, 2, seed=4), _ids(cfg, 2, seed=5)
    with torch.no_grad():
        oa = fast(a).last_hidden_state
        ob = fast(b).last_hidden_state
        wa, wb = ref(a), ref(b)
    assert oa.data_ptr() != ob.data_ptr()
    assert _rel(oa, wa.last_hidden_state) < 1e-2 and _rel(ob, wb.last_hidden_state) < 1e-2
    assert _rel(fast(b).pooler_output, wb.pooler_output) < 1e-2      # eos_token_id != 2: first-eos pooling
    assert len(fast.forward._cached) == 1
    with torch.no_grad():
        fast(_ids(cfg, 1, seed=6))
    assert len(fast.forward._cached) == 2


def test_in_place_weight_update_is_seen_by_the_next_call():
    cfg = _config(**TINY, act="quick_gelu", eos=2, vocab=1000)
    fast, ref = _build(cfg, False, torch.float16, seed=7)
    fast = _compile_pipe(text_encoder=fast)["text_encoder"]
    ids = _ids(cfg, 2, seed=8)
    with torch.no_grad():
        before = fast(ids).last_hidden_state
        w = fast.text_model.encoder.layers[1].mlp.fc2.weight
        delta = 0.05 * torch.randn_like(w)
        w.add_(delta)                                         # LoRA-style in-place merge
        ref.text_model.encoder.layers[1].mlp.fc2.weight.add_(delta.float())
        after = fast(ids).last_hidden_state
        want = ref(ids).last_hidden_state
    assert _rel(after, want) < 1e-2 and _rel(before, want) > 2e-2
    assert len(fast.forward._cached) == 1


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 1e-2), (torch.bfloat16, 4e-2)])
def test_sd15_text_encoder_full_size(dtype, tol):
    """CLIP ViT-L/14 text model (SD-1.5 / SDXL text_encoder): 12 layers, 768 wide, quick_gelu, legacy eos id."""
    cfg = _config(768, 12, 12, 3072, "quick_gelu", eos=2)
    fast, ref = _build(cfg, False, dtype, seed=9)
    assert sum(p.numel() for p in fast.parameters()) == 123_060_480
    eager16 = copy.deepcopy(fast)
    fast = _compile_pipe(text_encoder=fast)["text_encoder"]
    ids = _ids(cfg, 2, seed=10)
    with torch.no_grad():
        got, want, lib = fast(ids, output_hidden_states=True), ref(ids, output_hidden_states=True), eager16(ids)
    e = _rel(got.last_hidden_state, want.last_hidden_state)
    print(f"CLIP ViT-L/14 text {dtype}: native {e:.3e} | eager 16-bit {_rel(lib.last_hidden_state, want.last_hidden_state):.3e}")
    assert e < tol
    assert _rel(got.hidden_states[-2], want.hidden_states[-2]) < tol      # what SDXL / clip_skip read
    assert _rel(got.pooler_output, want.pooler_output) < tol


def test_sdxl_text_encoder_2_shape_with_projection():
    """OpenCLIP bigG text tower as SDXL's text_encoder_2 (1280 wide, 20 heads, gelu, text_projection);
    8 of its 32 layers keep the fp32 reference quick."""
    cfg = _config(1280, 20, 8, 5120, "gelu", eos=2, proj=1280)
    fast, ref = _build(cfg, True, torch.float16, seed=11)
    mods = _compile_pipe(text_encoder_2=fast)
    fast = mods["text_encoder_2"]
    ids = _ids(cfg, 2, seed=12)
    with torch.no_grad():
        got, want = fast(ids, output_hidden_states=True), ref(ids, output_hidden_states=True)
    assert type(got).__name__ == "CLIPTextModelOutput"
    assert _rel(got.text_embeds, want.text_embeds) < 1e-2
    assert _rel(got[0], want[0]) < 1e-2                                   # pipelines take [0] = text_embeds
    assert _rel(got.hidden_states[-2], want.hidden_states[-2]) < 1e-2
    assert _rel(got.last_hidden_state, want.last_hidden_state) < 1e-2


def test_compile_wraps_the_encoders_of_a_pipeline_object():
    from types import SimpleNamespace
    from sfast.compilers.diffusion_pipeline_compiler import CompilationConfig, compile
    from sfast_b200 import synthetic
    cfg = _config(**TINY, act="quick_gelu", eos=2, vocab=1000)
    te, ref = _build(cfg, False, torch.float16, seed=13)
    unet = synthetic.SyntheticUNet(synthetic.TINY, dtype=torch.float16, device="cuda")
    pipe = SimpleNamespace(unet=unet, text_encoder=te, device=torch.device("cuda"))
    c = CompilationConfig.Default()
    c.enable_cuda_graph = True
    pipe = compile(pipe, c)
    assert hasattr(pipe.text_encoder.forward, "_compiled")
    ids = _ids(cfg, 2, seed=14)
    with torch.no_grad():
        assert _rel(pipe.text_encoder(ids)[0], ref(ids)[0]) < 1e-2


def test_unsupported_calls_stay_on_the_modules_own_path():
    cfg = _config(**TINY, act="quick_gelu", eos=2, vocab=1000)
    fast, ref = _build(cfg, False, torch.float16, seed=15)
    fast = _compile_pipe(text_encoder=fast)["text_encoder"]
    ids = _ids(cfg, 2, seed=16)
    mask = torch.ones_like(ids)
    with torch.no_grad():
        got = fast(ids, attention_mask=mask)       # whole-module routing, logged once
        want = ref(ids, attention_mask=mask)
    assert _rel(got.last_hidden_state, want.last_hidden_state) < 1e-2
    assert len(fast.forward._cached) == 0


# ---------------------------------------------------------------------------------------------
# CLIP vision tower (SVD image_encoder)
# ---------------------------------------------------------------------------------------------
def _vision(hidden, heads, layers, inter, act, image, proj, dtype, seed, with_projection=True):
    cfg = transformers.CLIPVisionConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                                        num_attention_heads=heads, image_size=image, patch_size=14,
                                        projection_dim=proj, hidden_act=act)
    torch.manual_seed(seed)
    cls = transformers.CLIPVisionModelWithProjection if with_projection else transformers.CLIPVisionModel
    m = cls(cfg).eval()
    with torch.no_grad():
        for name, p in m.named_parameters():
            if "norm" in name:
                p.copy_(1.0 + 0.3 * torch.randn_like(p) if name.endswith("weight") else 0.3 * torch.randn_like(p))
            elif name.endswith(".bias"):
                p.copy_(0.1 * torch.randn_like(p))
            elif "class_embedding" in name or "position_embedding" in name:
                p.copy_(0.5 * torch.randn_like(p))
    ref = copy.deepcopy(m).to("cuda", torch.float32)
    return cfg, m.to("cuda", dtype), ref


def _compile_vision(m):
    from sfast.compilers.diffusion_pipeline_compiler import CompilationConfig, compile_image_encoder
    c = CompilationConfig.Default()
    c.enable_cuda_graph = True
    return compile_image_encoder(m, c)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 1e-2), (torch.bfloat16, 4e-2)])
def test_tiny_vision_tower_all_outputs(dtype, tol):
    cfg, fast, ref = _vision(128, 2, 2, 256, "quick_gelu", 56, 64, dtype, seed=21)
    eager16 = copy.deepcopy(fast)
    fast = _compile_vision(fast)
    x = torch.randn(3, 3, 56, 56, generator=torch.Generator().manual_seed(22)).to("cuda", dtype)
    with torch.no_grad():
        got = fast(x, output_hidden_states=True)
        want = ref(x.float(), output_hidden_states=True)
        lib = eager16(x)
    assert type(got).__name__ == "CLIPVisionModelOutput"
    e = _rel(got.image_embeds, want.image_embeds)
    print(f"tiny CLIP vision {dtype}: native {e:.3e} | eager 16-bit {_rel(lib.image_embeds, want.image_embeds):.3e}")
    assert e < tol
    assert _rel(got.last_hidden_state, want.last_hidden_state) < tol
    assert len(got.hidden_states) == len(want.hidden_states) == 3
    for g, w in zip(got.hidden_states, want.hidden_states):
        assert _rel(g, w) < tol
    x2 = torch.randn(3, 3, 56, 56, generator=torch.Generator().manual_seed(23)).to("cuda", dtype)
    with torch.no_grad():
        assert _rel(fast(x2).image_embeds, ref(x2.float()).image_embeds) < tol      # replay with new pixels
    assert len(fast.forward._cached) == 1


def test_vision_tower_without_projection_and_wrong_size():
    cfg, fast, ref = _vision(128, 2, 2, 256, "gelu", 56, 64, torch.float16, seed=24, with_projection=False)
    fast = _compile_vision(fast)
    x = torch.randn(2, 3, 56, 56, generator=torch.Generator().manual_seed(25)).to("cuda", torch.float16)
    with torch.no_grad():
        got, want = fast(x), ref(x.float())
    assert _rel(got.pooler_output, want.pooler_output) < 1e-2
    assert _rel(got[0], want[0]) < 1e-2
    with pytest.raises(ValueError):
        fast(torch.zeros(1, 3, 42, 42, device="cuda", dtype=torch.float16))


def test_svd_image_encoder_shape_vit_h14():
    """OpenCLIP ViT-H/14 vision tower as the SVD image_encoder: 1280 wide, 16 heads (head_dim 80), gelu,
    224^2 images -> 257 tokens, projection to 1024; 6 of its 32 layers keep the fp32 reference quick."""
    cfg, fast, ref = _vision(1280, 16, 6, 5120, "gelu", 224, 1024, torch.float16, seed=26)
    fast = _compile_vision(fast)
    x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(27)).to("cuda", torch.float16)
    with torch.no_grad():
        got, want = fast(x), ref(x.float())
    e = _rel(got.image_embeds, want.image_embeds)
    print(f"ViT-H/14 vision (6 layers) fp16: image_embeds {e:.3e}, last_hidden_state {_rel(got.last_hidden_state, want.last_hidden_state):.3e}")
    assert got.image_embeds.shape == (1, 1024) and e < 1e-2
    assert _rel(got.last_hidden_state, want.last_hidden_state) < 1e-2
