"""Where does a B=2 UNet step go?  Captures sub-graphs holding only one family / one GEMM shape of
the plan's ops (same buffers, same order) and times their replay: per-kernel steady-state cost
INCLUDING launch gaps, which per-kernel profilers do not show."""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-fast_b200"))
from sfast_b200.plan import PackedWeights, UNetPlan  # noqa: E402
from sfast_b200.synthetic import CONFIGS  # noqa: E402
from sfast_b200.unet_spec import random_state_dict, spec_from_config  # noqa: E402


def time_graph(ops, iters=20):
    if not ops:
        return 0.0
    st = torch.cuda.current_stream()
    for op in ops:
        op.launch(st.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        s = torch.cuda.current_stream().cuda_stream
        for op in ops:
            op.launch(s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    spec = spec_from_config(CONFIGS["sd15"])
    sd = random_state_dict(spec, seed=0, dtype=torch.float16, device="cuda")
    pw = PackedWeights(spec, sd, torch.float16, "cuda")
    plan = UNetPlan(pw, batch, 64, 64)
    ops = plan.all_ops()
    fam = collections.OrderedDict()
    for op in ops:
        name = getattr(op.fn, "__name__", "other")
        key = name
        if name == "sfb_gemm":
            p = op.keep[0]
            key = f"gemm M={p.M} N={p.N} K={p.K} s={p.splits} conv={p.a_mode} epi={p.epi}"
        elif name == "sfb_attention":
            p = op.keep[0]
            key = f"attn S={p.seq_q} Skv={p.seq_kv} D={p.head_dim}"
        fam.setdefault(key, []).append(op)
    total = time_graph([op for op in ops])
    print(json.dumps({"family": "ALL", "n": len(ops), "us": round(total, 1)}), flush=True)
    rows = []
    for key, lst in fam.items():
        us = time_graph(lst)
        rows.append((us, key, len(lst)))
    acc = 0.0
    for us, key, n in sorted(rows, reverse=True):
        acc += us
        print(json.dumps({"family": key, "n": n, "us": round(us, 1), "us_per_op": round(us / n, 2)}), flush=True)
    print(json.dumps({"family": "SUM_OF_PARTS", "us": round(acc, 1)}), flush=True)
    allg = [op for op in ops if getattr(op.fn, "__name__", "") == "sfb_gemm"]
    print(json.dumps({"family": "all gemm", "n": len(allg), "us": round(time_graph(allg), 1)}), flush=True)


if __name__ == "__main__":
    main()
