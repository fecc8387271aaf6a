"""Latency anatomy of one small tcgen05 GEMM launch: %globaltimer stamps from inside the kernel
(entry / setup / first TMA / first data / MMAs issued / accumulator ready / stored / exit), taken
on the last launch of a back-to-back chain inside a CUDA graph."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-fast_b200"))
from sfast_b200 import _lib, ops  # noqa: E402


def run(M, N, K, chain=8, residual=True, conv=None, rowbias=False, distinct=False):
    """conv = (n, h, w, cin): 3x3 implicit-GEMM convolution (M = n*h*w, K = 9*cin) instead of a
    plain matrix A operand.  distinct: every launch of the chain gets its own weights / input."""
    lib = _lib.lib()
    dt = torch.float16
    nbuf = chain if distinct else 1
    w = [torch.randn(N, K, device="cuda").to(dt) for _ in range(nbuf)]
    r = torch.randn(M, N, device="cuda").to(dt) if residual else None
    b = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=dt)
    mats = [ops.Mat(x) for x in w]
    mt, nt = (M + 127) // 128, (N + 159) // 160
    stamps = torch.zeros(chain, mt * nt, 8, dtype=torch.int64, device="cuda")
    oplist = []
    if conv:
        n, h, wd, cin = conv
        xs = [torch.randn(n, h, wd, cin, device="cuda").to(dt) for _ in range(nbuf)]
        box_n, box_h, box_w = ops.conv_tile_box(h, wd)
        rb = torch.randn(n, N, device="cuda") if rowbias else None
    else:
        a = [torch.randn(M, K, device="cuda").to(dt) for _ in range(nbuf)]
    for i in range(chain):
        j = i % nbuf
        if conv:
            ad = ops.a_conv(xs[j].data_ptr(), n, h, wd, cin, cin, box_n, box_h, box_w, 1)
            op = ops.gemm_op("c", lib, a=ad, b=mats[j], M=M, N=N, K=K, dt=dt, out=out, ldo=N, bias=b,
                             rowbias=rb, rows_per_img=h * wd, ld_rowbias=N, residual=r, ldr=N,
                             conv=dict(n=n, h=h, w=wd, cin=cin, stride=1, box_n=box_n, box_h=box_h, box_w=box_w))
            op.keep = tuple(op.keep) + (xs[j],)
        else:
            op = ops.gemm_op("g", lib, a=ops.a_matrix(a[j].data_ptr(), M, K, K), b=mats[j], M=M, N=N, K=K,
                             dt=dt, out=out, ldo=N, bias=b, residual=r, ldr=N)
            op.keep = tuple(op.keep) + (a[j],)
        op.keep[0].debug_stamps = stamps[i].data_ptr()
        oplist.append(op)
    st = torch.cuda.current_stream()
    for op in oplist:
        op.launch(st.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        s = torch.cuda.current_stream().cuda_stream
        for op in oplist:
            op.launch(s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    t = stamps.cpu().numpy().astype("int64")
    names = ["entry", "setup", "tma0", "data0", "mma_issued", "acc_ready", "stored", "exit"]
    res = {"M": M, "N": N, "K": K, "ctas": mt * nt, "conv": conv, "distinct": distinct}
    for i in (chain - 2, chain - 1):
        k = t[i]
        t0 = k[:, 0].min()
        prev_end = t[i - 1][:, 7].max()
        row = {"gap_prev_exit_to_first_entry_ns": int(t0 - prev_end),
               "last_entry_ns": int(k[:, 0].max() - t0)}
        for j, nm in enumerate(names[1:], 1):
            v = k[:, j]
            row[nm + "_med"] = int(sorted(v - k[:, 0])[len(v) // 2])
        row["total_first_entry_to_last_exit_ns"] = int(k[:, 7].max() - t0)
        res[f"launch{i}"] = row
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    run(8192, 320, 320)
    run(2048, 640, 640)
    run(512, 1280, 1280)
    run(8192, 320, 2880)
    run(8192, 320, 2880, distinct=True)
    run(8192, 320, 2880, conv=(2, 64, 64, 320))
    run(8192, 320, 2880, conv=(2, 64, 64, 320), rowbias=True, distinct=True)
    run(2048, 640, 5760, conv=(2, 32, 32, 640), rowbias=True, distinct=True)
