"""Generates tests/golden/tiny_unet_fp32.pt from the oracle (CPU, fp32, seeded).

The real reference cannot be imported in this image (`import sfast` hard-requires the compiled
`sfast._C`, /root/reference/src/sfast/__init__.py:22-31, and diffusers/xformers are absent), so the
golden vector is produced by the oracle restatement: UNet-level parity is "unpinned" against the
reference itself (see oracle/unet_oracle.py header).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import unet_oracle as uo  # noqa: E402


def main():
    seed = 1234
    cfg = uo.tiny_config()
    m = uo.build_unet(cfg, seed=seed)
    g = torch.Generator().manual_seed(99)
    sample = torch.randn(2, 4, 32, 32, generator=g)
    ehs = torch.randn(2, 77, cfg.cross_attention_dim, generator=g)
    t = torch.tensor(801.0)
    with torch.no_grad():
        out = m(sample, t, ehs).sample
    chk = sum(float(p.double().abs().sum()) for p in m.parameters())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_unet_fp32.pt")
    torch.save({"seed": seed, "sample": sample, "timestep": t, "encoder_hidden_states": ehs,
                "out": out, "param_abs_sum": chk}, path)
    print("wrote", path, out.shape, float(out.std()))


if __name__ == "__main__":
    main()
