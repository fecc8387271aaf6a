"""GPU parity of the VAE decoder (`compile_vae` -> AutoencoderKL.decode on the native path) against
oracle/vae_oracle.py on the same seeded weights (random norm affines) and synthetic latents."""
import pytest
import torch

from oracle import vae_oracle as vo

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _rel(got, ref):
    got, ref = got.float(), ref.float()
    d = (got - ref).abs()
    rms = ref.pow(2).mean().sqrt()
    return max((d.max() / ref.abs().max()).item(), (d / (ref.abs() + rms)).max().item())


def _compile(m, graph=True):
    from sfast.compilers.diffusion_pipeline_compiler import CompilationConfig, compile_vae
    c = CompilationConfig.Default()
    c.enable_cuda_graph = graph
    return compile_vae(m, c)


def _pair(cfg, seed, dtype=torch.float16):
    oracle = vo.build_vae(cfg, seed=seed, dtype=torch.float32, device="cuda")
    fast = vo.build_vae(cfg, seed=seed, dtype=dtype, device="cuda")
    oracle.load_state_dict({k: v.float() for k, v in fast.state_dict().items()})
    return oracle, fast


@pytest.mark.parametrize("graph", [False, True])
def test_tiny_vae_decode_vs_oracle(graph):
    cfg = vo.tiny_vae_config()
    oracle, fast = _pair(cfg, seed=4)
    fast = _compile(fast, graph)
    for b, h, w in ((1, 16, 16), (2, 16, 24), (1, 32, 32)):
        g = torch.Generator(device="cuda").manual_seed(b * 100 + h)
        z = torch.randn(b, 4, h, w, device="cuda", generator=g).half()
        got = fast.decode(z).sample
        with torch.no_grad():
            ref = oracle.decode(z.float()).sample
        assert got.shape == ref.shape == (b, 3, 8 * h, 8 * w)
        err = _rel(got, ref)
        print(f"tiny VAE {b}x{h}x{w}: rel err {err:.3e}")
        assert err < TOL


@pytest.mark.parametrize("dtype,tol", [(torch.float16, TOL), (torch.bfloat16, 4e-2)])
def test_sd_vae_decode_512(dtype, tol):
    """The SD-1.5 VAE decoder at its pipeline shape: 4 x 64 x 64 latent -> 3 x 512 x 512 image."""
    cfg = vo.sd_vae_config()
    oracle, fast = _pair(cfg, seed=1, dtype=dtype)
    fast = _compile(fast, True)
    g = torch.Generator(device="cuda").manual_seed(9)
    z = (torch.randn(1, 4, 64, 64, device="cuda", generator=g) / 0.18215 * 0.2).to(dtype)
    got = fast.decode(z, return_dict=False)[0]
    with torch.no_grad():
        ref = oracle.decode(z.float()).sample
    err = _rel(got, ref)
    print(f"SD VAE decode 512x512 {dtype}: rel err {err:.3e}")
    assert err < tol
    assert len(fast.decode._cached) == 1
