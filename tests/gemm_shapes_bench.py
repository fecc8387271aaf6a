"""Per-shape timing of sfb_gemm (legacy one-tile kernel vs the persistent pair kernel):
    python tests/gemm_shapes_bench.py [tag]        # one JSON line per (shape, kernel)
Each shape is captured as a CUDA graph of 20 back-to-back launches and replayed (CUDA events).
SFB_LIB_PATH selects an alternative build of the library (A/B of kernel variants)."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-fast_b200"))
from sfast_b200 import _lib, ops  # noqa: E402
from sfast_b200.ops import Act  # noqa: E402

DEV = "cuda"
# (name, kind, args): shapes of SD-1.5 at B = 8 (64^2 / 128^2 latents) and SDXL at B = 8 (128^2)
SHAPES = [
    ("conv 320->320 @64^2 B8", "conv", dict(n=8, h=64, w=64, cin=320, cout=320)),
    ("conv 640->640 @32^2 B8", "conv", dict(n=8, h=32, w=32, cin=640, cout=640)),
    ("conv 1280->1280 @16^2 B8", "conv", dict(n=8, h=16, w=16, cin=1280, cout=1280)),
    ("conv 320->320 @128^2 B8", "conv", dict(n=8, h=128, w=128, cin=320, cout=320)),
    ("conv 960->320 @64^2 B8", "conv", dict(n=8, h=64, w=64, cin=960, cout=320)),
    ("linear M32768 N320 K320", "gemm", dict(M=32768, N=320, K=320)),
    ("linear M32768 N960 K320 (qkv)", "gemm", dict(M=32768, N=960, K=320)),
    ("linear M32768 N320 K1280 (ff.out)", "gemm", dict(M=32768, N=320, K=1280)),
    ("geglu M32768 K320 inner1280", "geglu", dict(M=32768, K=320, inner=1280)),
    ("geglu M8192 K640 inner2560", "geglu", dict(M=8192, K=640, inner=2560)),
    ("sdxl linear M32768 N640 K640", "gemm", dict(M=32768, N=640, K=640)),
    ("sdxl linear M8192 N1280 K1280", "gemm", dict(M=8192, N=1280, K=1280)),
    ("sdxl geglu M8192 K1280 inner5120", "geglu", dict(M=8192, K=1280, inner=5120)),
    ("sdxl ff.out M8192 N1280 K5120", "gemm", dict(M=8192, N=1280, K=5120)),
    ("gemm 8192^2 x 1280 (square-ish)", "gemm", dict(M=8192, N=8000, K=1280)),
]


def build(kind, a, persistent, dt=torch.float16):
    lib = _lib.lib()
    keep = []
    if kind == "conv":
        n, h, w, cin, cout = a["n"], a["h"], a["w"], a["cin"], a["cout"]
        x = torch.randn(n, h, w, cin, device=DEV).to(dt)
        wt = (torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(9 * cin)).to(dt)
        b = torch.randn(cout, device=DEV)
        M = n * h * w
        out = torch.zeros(M, cout, device=DEV, dtype=dt)
        res = torch.randn(M, cout, device=DEV).to(dt)
        bn, bh, bw = ops.conv_tile_box(h, w)
        op = ops.gemm_op("conv", lib, a=ops.a_conv(x.data_ptr(), n, h, w, cin, cin, bn, bh, bw, 1),
                         b=ops.Mat(ops.pack_conv3x3(wt, dt)), M=M, N=cout, K=9 * cin, dt=dt, out=out, ldo=cout,
                         bias=b, residual=res, ldr=cout, splits=1, persistent=persistent,
                         conv=dict(n=n, h=h, w=w, cin=cin, stride=1, box_n=bn, box_h=bh, box_w=bw))
        keep += [x, wt, b, out, res]
    elif kind == "gemm":
        M, N, K = a["M"], a["N"], a["K"]
        x = torch.randn(M, K, device=DEV).to(dt)
        w = (torch.randn(N, K, device=DEV) / math.sqrt(K)).to(dt)
        b = torch.randn(N, device=DEV)
        out = torch.zeros(M, N, device=DEV, dtype=dt)
        op = ops.gemm_op("gemm", lib, a=ops.a_matrix(x.data_ptr(), M, K, K), b=ops.Mat(w), M=M, N=N, K=K, dt=dt,
                         out=out, ldo=N, bias=b, residual=out, ldr=N, splits=1, persistent=persistent)
        keep += [x, w, b, out]
    else:
        M, K, inner = a["M"], a["K"], a["inner"]
        x = torch.randn(M, K, device=DEV).to(dt)
        w = (torch.randn(2 * inner, K, device=DEV) / math.sqrt(K)).to(dt)
        b = torch.randn(2 * inner, device=DEV) * 0.1
        wp, bp, _ = ops.pack_geglu(w, b, dt)
        out = torch.zeros(M, inner, device=DEV, dtype=dt)
        op = ops.gemm_op("geglu", lib, a=ops.a_matrix(x.data_ptr(), M, K, K), b=ops.Mat(wp), M=M, N=wp.shape[0],
                         K=K, dt=dt, out=out, ldo=inner, bias=bp, epi=ops.EPI_GEGLU, geglu_n_out=inner,
                         splits=1, persistent=persistent)
        keep += [x, w, wp, bp, out]
    return op, keep


def time_op(op, reps=20, iters=5):
    st = torch.cuda.current_stream()
    op.launch(st.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cs = torch.cuda.current_stream().cuda_stream
        for _ in range(reps):
            op.launch(cs)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)  # us per launch


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("SFB_LIB_PATH", "default")
    only = os.environ.get("SFB_SHAPES")  # substring filter (profiling one shape under ncu)
    for name, kind, a in SHAPES:
        if only and only not in name:
            continue
        row = {"lib": os.path.basename(tag), "shape": name}
        for label, persistent in (("legacy", False), ("persistent", True)):
            try:
                op, keep = build(kind, a, persistent)
                us = time_op(op)
                row[label + "_us"] = round(us, 2)
                row[label + "_tflops"] = round(op.flops / us / 1e6, 1)
                del op, keep
            except Exception as exc:  # noqa: BLE001
                row[label + "_error"] = repr(exc)[:200]
            torch.cuda.empty_cache()
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
