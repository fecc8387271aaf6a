"""Times the attention kernel alone (CUDA events, L2-flushing rotation of buffers) for the shapes
of SD-1.5 / SDXL / SVD.  SFB_LIB_PATH selects an alternative build of the library (A/B of the
MUFU / polynomial exponential split, -DSFB_EXP_POLY_MASK=...)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import kernel_checks as kc  # noqa: E402
from sfast_b200 import _lib, ops  # noqa: E402


def bench(B, H, S, Skv, D, iters=20):
    lib = _lib.lib()
    dt = torch.float16
    sets = []
    for _ in range(3):
        q, k, vt, dv, q_pitch, vt_pitch = kc._attn_buffers(B, H, S, Skv, D, dt)
        q.normal_(); k.normal_()
        vt.view(B * H, dv, vt_pitch)[:, :D, :Skv].normal_()
        out = torch.zeros(B, S, H * D, device="cuda", dtype=dt)
        sets.append(ops.attention_op("a", lib, q=q, k=k, vt=vt, out=out, batch=B, heads=H, head_dim=D,
                                     seq_q=S, seq_kv=Skv, q_rows=S, k_rows=Skv, vt_rows=dv,
                                     q_pitch=q_pitch, vt_pitch=vt_pitch, dt=dt))
    st = torch.cuda.current_stream().cuda_stream
    for op in sets:
        op.launch(st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        sets[i % 3].launch(st)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    fl = 4 * B * H * S * Skv * D
    print(json.dumps({"lib": os.path.basename(os.environ.get("SFB_LIB_PATH", "libsfb200.so")), "B": B, "S": S, "Skv": Skv, "D": D,
                      "us": round(us, 1), "tflops": round(fl / us / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    bench(2, 8, 4096, 4096, 40)
    bench(16, 8, 4096, 4096, 40)
    bench(2, 8, 1024, 1024, 80)
    bench(2, 8, 4096, 77, 40)
    bench(8, 10, 4096, 4096, 64)
    bench(8, 8, 16384, 16384, 40, iters=4)
    bench(10, 5, 9216, 9216, 64, iters=6)
