"""Latency anatomy of one small tcgen05 GEMM launch: %globaltimer stamps from inside the kernel
(entry / setup / PDL wait passed / first data / MMAs issued / accumulator ready / staged / stored /
synced / exit), taken on the last launches of a back-to-back chain inside a CUDA graph.

Needs the measurement build of the library (the product build carries no instrumentation):
    cd stable-fast_b200/csrc && make trace        # -> sfast_b200/libsfb200_trace.so
    SFB_LIB_PATH=stable-fast_b200/sfast_b200/libsfb200_trace.so python tests/gemm_latency.py"""
import ctypes
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
    stamps = torch.zeros(chain, mt * nt, 16, dtype=torch.int64, device="cuda")
    raw = ctypes.CDLL(_lib.LIB_PATH)
    raw.sfb_trace_next_gemm.argtypes = [ctypes.c_void_p]
    raw.sfb_trace_next_gemm.restype = None
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
        oplist.append(op)
    st = torch.cuda.current_stream()
    for op in oplist:
        op.launch(st.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        s = torch.cuda.current_stream().cuda_stream
        for i, op in enumerate(oplist):
            raw.sfb_trace_next_gemm(stamps[i].data_ptr())
            op.launch(s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    t = stamps.cpu().numpy().astype("int64")
    names = ["entry", "setup", "pdl_released", "data0", "mma_issued", "acc_ready", "staged", "stored_first_thread",
             "synced", "exit", "stored_last_thread"]
    res = {"M": M, "N": N, "K": K, "ctas": mt * nt, "conv": conv, "distinct": distinct, "residual": residual}
    for i in (chain - 2, chain - 1):
        k = t[i]
        k = k[k[:, 0] > 0]                      # (pair grids: every CTA stamps)
        prev = t[i - 1]
        prev = prev[prev[:, 0] > 0]
        # all times relative to the PREVIOUS launch's last exit = the moment this launch's inputs exist
        prev_end = prev[:, 9].max()
        row = {"first_entry_vs_prev_exit_ns": int(k[:, 0].min() - prev_end)}
        for j, nm in enumerate(names):
            v = k[:, j]
            v = v[v > 0]
            if len(v):
                row[nm + "_med"] = int(sorted(v - prev_end)[len(v) // 2])
        row["last_exit_ns"] = int(k[:, 9].max() - prev_end)
        res[f"launch{i}"] = row
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    run(8192, 320, 320)
    run(8192, 320, 320, residual=False)
    run(2048, 640, 640)
    run(512, 1280, 1280)
    run(8192, 320, 2880)
    run(8192, 320, 2880, distinct=True)
    run(8192, 320, 2880, conv=(2, 64, 64, 320))
    run(8192, 320, 2880, conv=(2, 64, 64, 320), rowbias=True, distinct=True)
    run(2048, 640, 5760, conv=(2, 32, 32, 640), rowbias=True, distinct=True)
