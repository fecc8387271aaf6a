"""Pins against outputs of the REFERENCE's own kernels.

tests/golden/ref_triton_norms.pt holds what /root/reference/src/sfast/triton/ops/group_norm.py and
layer_norm.py (unmodified, imported from /root/reference) computed under Triton's interpreter in
the build container -- see tests/golden/make_reference_golden.py.  Rows a4 (GroupNorm / GroupNorm +
SiLU) and a9 (LayerNorm) of SURVEY.md section 8 are therefore pinned to reference-run vectors
(zero-mean cases at the reference's own 1e-2; the large-mean case with the reference's own fp16
arithmetic error added, see _stat_slack):

  * CPU (`-m "not gpu"`): oracle/ops_oracle.py against the reference outputs;
  * GPU (`-m gpu`): the CUDA kernels (fused and two-pass GroupNorm, stand-alone LayerNorm, and
    the LayerNorm folded into a tcgen05 GEMM) through the C ABI against the same outputs.

Tolerance = the reference's own self-test: assert_close(rtol=1e-2, atol=1e-2)
(group_norm.py:499,523; layer_norm.py:435 uses atol=1e-2, rtol=0).  Note the reference round-trips
mean / rstd through fp16 (group_norm.py:405-416), so ITS outputs carry ~1e-2 of error against an
fp32 evaluation; the CUDA path keeps the statistics in fp32 and is compared with that slack.
"""
import math
import os

import pytest
import torch

from oracle import ops_oracle as oo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_triton_norms.pt")


def _fx():
    return torch.load(GOLDEN)


def _close(got, ref, rtol=1e-2, atol=1e-2):
    torch.testing.assert_close(got.float().cpu(), ref.float().cpu(), rtol=rtol, atol=atol)


def _stat_slack(case, key):
    """Extra absolute slack for a reference OUTPUT on inputs with |mean| >> sigma.  The reference's
    channels_last apply kernel works in the INPUT dtype end to end (group_norm.py:295-318: mean,
    rstd, gamma, beta are fp16, `a = rstd * gamma`, `b = beta - a * mean`, `x = a * x + b` all in
    fp16) after storing mean / rstd in fp16 (group_norm.py:405-416).  So its own result is off by
    about |a| * (|mean - mean_fp16| + 2^-10 * |mean|): ~0 for zero-mean data, ~0.15 when the data
    sit at 50 +- 1.  The CUDA path and the oracle compute in fp32; this is the reference's error,
    documented here, not a tolerance the product path needs."""
    n = case["x"].shape[0]
    xf = case["x"].float().reshape(n, case["groups"], -1)
    mean = xf.mean(-1)
    dmean = (mean - case["mean_" + key].float()).abs().max()
    a = case["weight"].float().abs().max() * case["rstd_" + key].float().max()
    return float(a * (dmean + 2.0 ** -10 * mean.abs().max()))


def _with_reference_stats(case, key, silu):
    """The op evaluated in fp32 from the reference's OWN stored (fp16) statistics: isolates the
    arithmetic of the apply step from the statistics round trip."""
    x = case["x"].float()
    n, c = x.shape[:2]
    cpg = c // case["groups"]
    mean = case["mean_" + key].float().repeat_interleave(cpg, 1)[:, :, None, None]
    rstd = case["rstd_" + key].float().repeat_interleave(cpg, 1)[:, :, None, None]
    y = (x - mean) * rstd * case["weight"].float()[None, :, None, None] + case["bias"].float()[None, :, None, None]
    return y * torch.sigmoid(y) if silu else y


def test_fixture_was_produced_by_the_reference_files():
    fx = _fx()
    assert fx["generator"].endswith("make_reference_golden.py")
    assert "triton/ops/group_norm.py" in fx["reference_files"]
    assert len(fx["group_norm"]) >= 4 and len(fx["layer_norm"]) >= 3
    tags = [c["tag"] for c in fx["group_norm"]]
    assert "selftest_2x320x32x32" in tags  # the reference self-test configuration


@pytest.mark.parametrize("silu", [False, True])
def test_ops_oracle_group_norm_matches_reference_triton_outputs(silu):
    key = "silu" if silu else "plain"
    for case in _fx()["group_norm"]:
        got = oo.group_norm(case["x"], case["groups"], case["weight"], case["bias"], case["eps"], silu)
        slack = _stat_slack(case, key)
        _close(got, case["y_nhwc_" + key], atol=1e-2 + slack)
        _close(_with_reference_stats(case, key, silu), case["y_nhwc_" + key], atol=1e-2 + slack)
        if "y_nchw_" + key in case:
            _close(got, case["y_nchw_" + key], atol=1e-2 + slack)
        # the statistics themselves (the reference stores them in fp16)
        n, c = case["x"].shape[:2]
        xf = case["x"].float().reshape(n, case["groups"], -1)
        _close(xf.mean(-1), case["mean_" + key], rtol=1e-3, atol=2e-3)  # fp16 storage of the mean
        _close(1.0 / torch.sqrt(xf.var(-1, unbiased=False) + case["eps"]), case["rstd_" + key],
               rtol=2e-3, atol=2e-3)


def test_ops_oracle_layer_norm_matches_reference_triton_outputs():
    for case in _fx()["layer_norm"]:
        got = oo.layer_norm(case["x"], case["weight"], case["bias"], case["eps"])
        _close(got, case["y"], rtol=0, atol=1e-2)


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("silu", [False, True])
def test_cuda_group_norm_matches_reference_triton_outputs(silu, fused):
    from sfast_b200 import _lib, ops
    from sfast_b200.ops import Act
    lib = _lib.lib()
    key = "silu" if silu else "plain"
    for case in _fx()["group_norm"]:
        x = case["x"].cuda()
        n, c, h, w = x.shape
        xb = x.permute(0, 2, 3, 1).contiguous()                      # the path's NHWC layout
        yb = torch.zeros_like(xb)
        stats = torch.empty(ops.gn_ws_floats(n, case["groups"]), device="cuda")
        sync = torch.zeros(4, device="cuda", dtype=torch.int32)
        gops = ops.gn_ops("gn", lib, x=Act(xb, n, h, w, c), y=Act(yb, n, h, w, c),
                          gamma=case["weight"].cuda().float(), beta=case["bias"].cuda().float(),
                          stats=stats, groups=case["groups"], eps=case["eps"], silu=silu,
                          dt=torch.float16, sync=sync if fused else None)
        assert len(gops) == (1 if fused else 2)
        for op in gops:
            op.launch(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        _close(yb.permute(0, 3, 1, 2), case["y_nhwc_" + key], atol=1e-2 + _stat_slack(case, key))
        # and tightly against the fp32-statistics oracle (what the reference approximates)
        want = oo.group_norm(case["x"], case["groups"], case["weight"], case["bias"], case["eps"], silu)
        _close(yb.permute(0, 3, 1, 2), want, rtol=4e-3, atol=4e-3)


@pytest.mark.gpu
def test_cuda_layer_norm_and_folded_layer_norm_match_reference_triton_outputs():
    from sfast_b200 import _lib, ops
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    for case in _fx()["layer_norm"]:
        x = case["x"].cuda()
        rows, c = x.shape
        gamma, beta = case["weight"].cuda().float(), case["bias"].cuda().float()
        # (1) stand-alone kernel
        y = torch.zeros_like(x)
        ops.ln_op("ln", lib, x=x, y=y, rows=rows, c=c, gamma=gamma, beta=beta, eps=case["eps"],
                  dt=torch.float16).launch(st)
        torch.cuda.synchronize()
        _close(y, case["y"], rtol=0, atol=1e-2)
        # (2) the product path: LayerNorm folded into the consuming tcgen05 GEMM.  Row statistics as
        # the producer epilogue would leave them (sum, sum of squares of the stored fp16 row); the
        # expected result is the REFERENCE LayerNorm output pushed through the same linear in fp32.
        torch.manual_seed(1)
        n_out = 320
        w = (torch.randn(n_out, c, device="cuda") / math.sqrt(c)).half()
        b = torch.randn(n_out, device="cuda") * 0.1
        wp, bias, colsum = ops.fold_layer_norm(w, b, gamma, beta, torch.float16)
        xf = x.float()
        stats = torch.stack([xf.sum(-1), (xf * xf).sum(-1)], dim=-1).contiguous()
        out = torch.zeros(rows, n_out, device="cuda", dtype=torch.float16)
        ops.gemm_op("consumer", lib, a=ops.a_matrix(x.data_ptr(), rows, c, c), b=ops.Mat(wp.contiguous()),
                    M=rows, N=n_out, K=c, dt=torch.float16, out=out, ldo=n_out, bias=bias, splits=1,
                    ln=dict(rowstats=stats, colsum=colsum, eps=case["eps"], dim=c)).launch(st)
        torch.cuda.synchronize()
        ref = case["y"].cuda().float() @ w.float().t() + b
        # the reference LN output is itself fp16-rounded: allow its quantisation through the GEMM
        _close(out, ref, rtol=1e-2, atol=2e-2)


# ------------------------------------------------------------------------------------------------
# Row a7: Linear + GEGLU pinned against the reference's own CUTLASS kernel RUN ON THE B200
# (/root/reference/src/sfast/csrc/operators/cutlass/cutlass_dual_linear_kernel.cu, compiled from its
# own source by oracle/build_ref.sh into oracle/_ref/ -- the git-ignored .so travels with the
# snapshot).
_REF_GEGLU = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                          "libsfast_ref_geglu.so")


def _ref_geglu_op():
    if not os.path.exists(_REF_GEGLU):
        pytest.skip("oracle/_ref/libsfast_ref_geglu.so not built (run oracle/build_ref.sh where "
                    "/root/reference exists)")
    torch.ops.load_library(_REF_GEGLU)
    return torch.ops.sfast_ref.linear_geglu


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(4608, 320, 1280), (1152, 640, 2560), (300, 1280, 5120), (18432, 320, 1280)])
def test_cuda_geglu_matches_reference_cutlass_kernel(dtype, shape):
    """UNet GEGLU sizes (the reference's benchmark test uses in_features 320 / 640 -> 1280 / 2560,
    tests/operators/test_cutlass_dual_linear.py:59-72) at the reference's tolerance 2e-2: the
    reference kernel's output, the oracle and the tcgen05 GEMM + GEGLU epilogue must agree.  The
    reference accumulates in fp16 / bf16 and uses the tanh GELU under torch's default flags
    (cutlass_dual_linear_kernel.cu:391-393,509-514); the CUDA path accumulates in fp32 with erf."""
    from sfast_b200 import _lib, ops
    ref_op = _ref_geglu_op()
    lib = _lib.lib()
    m, k, inner = shape
    torch.manual_seed(m + k)
    x = torch.randn(m, k, device="cuda").to(dtype)
    w = (torch.randn(2 * inner, k, device="cuda") / math.sqrt(k)).to(dtype)
    b = (torch.randn(2 * inner, device="cuda") * 0.1).to(dtype)
    ref = ref_op(x, w, b)
    assert ref.shape == (m, inner) and ref.dtype == dtype
    want = oo.linear_geglu(x, w, b)
    wp, bp, _ = ops.pack_geglu(w, b.float(), dtype)
    out = torch.zeros(m, inner, device="cuda", dtype=dtype)
    ops.gemm_op("geglu", lib, a=ops.a_matrix(x.data_ptr(), m, k, k), b=ops.Mat(wp), M=m, N=wp.shape[0],
                K=k, dt=dtype, out=out, ldo=inner, bias=bp, epi=ops.EPI_GEGLU, geglu_n_out=inner,
                splits=1).launch(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    tol = 2e-2 if dtype == torch.float16 else 6e-2   # the bf16 reference accumulates in bf16
    _close(want, ref, rtol=tol, atol=tol)             # oracle  vs reference-run
    _close(out, ref, rtol=tol, atol=tol)              # product vs reference-run
    _close(out, want, rtol=1e-2, atol=1e-2)           # product vs oracle (fp32 truth), tighter


@pytest.mark.gpu
@pytest.mark.parametrize("bias", [False, True])
def test_reference_cutlass_kernel_on_its_own_test_configuration(bias):
    """The reference's own small-shape test (tests/operators/test_cutlass_dual_linear.py:42-56) run
    against the oracle -- checks the oracle's chunk order / GELU against the reference binary."""
    ref_op = _ref_geglu_op()
    for n in (4, 16):
        for cin in (8, 16):
            for cout in (8, 16):
                torch.manual_seed(n * 100 + cin * 10 + cout)
                proj = torch.nn.Linear(cin, 2 * cout, bias=bias).cuda().half()
                x = torch.randn(n, cin, device="cuda").half()
                got = ref_op(x, proj.weight, proj.bias)
                _close(oo.linear_geglu(x, proj.weight, proj.bias), got, rtol=2e-2, atol=2e-2)
