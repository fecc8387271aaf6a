"""GroupNorm+SiLU -> conv3x3: separate GroupNorm kernel + 9-tap TMA conv vs statistics kernel + halo conv
(GroupNorm applied on the conv's operand path, SFB_A_CONV3X3_GN).
    python tests/conv_gn_bench.py      # one JSON line per shape: us per (norm + conv) pair
Each variant is captured as a CUDA graph of 10 back-to-back (norm, conv) pairs and replayed."""
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
SHAPES = [
    dict(n=2, h=64, w=64, cin=320, cout=320),
    dict(n=2, h=64, w=64, cin=960, cout=320),
    dict(n=2, h=32, w=32, cin=640, cout=640),
    dict(n=8, h=64, w=64, cin=320, cout=320),
    dict(n=8, h=64, w=64, cin=960, cout=320),
    dict(n=8, h=32, w=32, cin=640, cout=640),
    dict(n=8, h=32, w=32, cin=1920, cout=640),
    dict(n=8, h=16, w=16, cin=1280, cout=1280),
    dict(n=8, h=128, w=128, cin=320, cout=320),
    dict(n=1, h=512, w=512, cin=128, cout=128),   # VAE decoder tail
]


def build(a, folded, dt=torch.float16, groups=32):
    lib = _lib.lib()
    n, h, w, cin, cout = a["n"], a["h"], a["w"], a["cin"], a["cout"]
    xb = torch.randn(n, h, w, cin, device=DEV).to(dt)
    x = Act(xb, n, h, w, cin)
    gamma, beta = torch.randn(cin, device=DEV), torch.randn(cin, device=DEV)
    wt = (torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(9 * cin)).to(dt)
    wm = ops.Mat(ops.pack_conv3x3(wt, dt))
    b = torch.randn(cout, device=DEV)
    M = n * h * w
    out = torch.zeros(M, cout, device=DEV, dtype=dt)
    res = torch.randn(M, cout, device=DEV).to(dt)
    stats = torch.zeros(ops.gn_ws_floats(n, groups), device=DEV)
    keep = [xb, wt, b, out, res, stats, gamma, beta]
    if folded:
        cnt = torch.zeros(max(n, 4), device=DEV, dtype=torch.int32)
        ab = torch.zeros(n, cin, 2, device=DEV)
        o1 = [ops.gn_scale_shift_op("gn", lib, x=x, gamma=gamma, beta=beta, stats=stats, counters=cnt,
                                    scale_shift=ab, groups=groups, eps=1e-5, dt=dt)]
        o2 = ops.gemm_op("conv_gn", lib, a=ops.a_conv_halo(x.ptr, n, h, w, cin, cin), b=wm, M=M, N=cout, K=9 * cin,
                         dt=dt, out=out, ldo=cout, bias=b, residual=res, ldr=cout,
                         conv=dict(n=n, h=h, w=w, cin=cin, stride=1, box_n=1, box_h=ops.HALO_BOX_H, box_w=ops.HALO_BOX_W),
                         gn=dict(scale_shift=ab, silu=True))
        keep += [cnt, ab]
    else:
        yb = torch.zeros(n, h, w, cin, device=DEV, dtype=dt)
        y = Act(yb, n, h, w, cin)
        sync = torch.zeros(4, device=DEV, dtype=torch.int32)
        o1 = ops.gn_ops("gn", lib, x=x, y=y, gamma=gamma, beta=beta, stats=stats, groups=groups, eps=1e-5, silu=True,
                        dt=dt, sync=sync)
        bn, bh, bw = ops.conv_tile_box(h, w)
        o2 = ops.gemm_op("conv", lib, a=ops.a_conv(y.ptr, n, h, w, cin, cin, bn, bh, bw, 1), b=wm, M=M, N=cout,
                         K=9 * cin, dt=dt, out=out, ldo=cout, bias=b, residual=res, ldr=cout, splits=1,
                         conv=dict(n=n, h=h, w=w, cin=cin, stride=1, box_n=bn, box_h=bh, box_w=bw))
        keep += [yb, sync]
    return o1, o2, keep, (sync if not folded else None)


def time_pair(o1, o2, sync, reps=10, iters=5, conv_only=False):
    def run(cs):
        if not conv_only:
            if sync is not None:
                sync.zero_()
            for o in o1:
                o.launch(cs)
        o2.launch(cs)
    run(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cs = torch.cuda.current_stream().cuda_stream
        for _ in range(reps):
            run(cs)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)


def main():
    only = os.environ.get("SFB_SHAPES")
    for a in SHAPES:
        name = "n{n} {h}x{w} {cin}->{cout}".format(**a)
        if only and only not in name:
            continue
        row = {"shape": name}
        for label, folded in (("separate", False), ("folded", True)):
            if os.environ.get("SFB_ONLY") and os.environ["SFB_ONLY"] != label:   # one variant (ncu captures)
                continue
            try:
                o1, o2, keep, sync = build(a, folded)
                us = time_pair(o1, o2, sync)
                usc = time_pair(o1, o2, sync, conv_only=True)
                row[label + "_us"] = round(us, 2)
                row[label + "_conv_only_us"] = round(usc, 2)
                row[label + "_conv_tflops"] = round(o2.flops / usc / 1e6, 1)
                del o1, o2, keep
            except Exception as exc:  # noqa: BLE001
                row[label + "_error"] = repr(exc)[:300]
            torch.cuda.empty_cache()
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
