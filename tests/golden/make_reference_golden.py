"""Generates tests/golden/ref_triton_norms.pt: outputs of the REFERENCE's own Triton kernels.

What runs: the unmodified files /root/reference/src/sfast/triton/ops/{group_norm,layer_norm,
activation,utils}.py (+ sfast/utils/copy_func.py), imported from where they lie -- nothing is
copied.  `import sfast` itself is impossible here (src/sfast/__init__.py:22-31 needs the compiled
sfast._C), so the package objects `sfast`, `sfast.triton`, `sfast.triton.ops`, `sfast.utils` are
registered as empty namespace modules whose __path__ points into /root/reference: the four kernel
files then import exactly as in the reference.  There is no GPU in this container, so the
@triton.jit kernels execute under Triton's own interpreter (TRITON_INTERPRET=1, numpy on the CPU):
same kernel source, same launch grids, same host wrappers (`group_norm_forward`,
`group_norm_silu_forward`, `LayerNorm.forward`).  Two harness shims, neither touches reference
code: (1) the script must run under `python -O` so the wrappers' `assert input.device.type ==
'cuda'` (group_norm.py:390) is skipped; (2) Triton 3.6's interpreter looks a kernel up by
`fn.__name__` after re-exec'ing its source, which breaks for the reference's renamed copies
(`copy_func(..., name=f'{kernel.__name__}_{act.__name__}')`, group_norm.py:87-96) -- the lookup
falls back to the single function the source defines.

Cases: the reference self-test configuration (group_norm.py:481-527: randn(2, 320, 32, 32) fp16,
G = 32, randn affine, eps 1e-5, NCHW and channels_last, with and without SiLU), UNet-shaped
extras (1280 channels at 8x8, a large-mean input, an outlier channel), and LayerNorm on the
self-test's data recipe (layer_norm.py:406-438: x = -2.3 + 0.5 randn, rand affine) at UNet widths.

Run from the repo root (about 5 minutes):   python -O tests/golden/make_reference_golden.py
"""
import importlib
import os
import sys
import types

os.environ["TRITON_INTERPRET"] = "1"

import torch  # noqa: E402

REF = "/root/reference/src/sfast"


def _load_reference_kernels():
    if __debug__:
        raise SystemExit("run with `python -O` (the reference wrappers assert a CUDA device)")
    from triton.runtime import interpreter as ti

    def compile_and_exec(self, tree):
        code = compile(tree, filename=self.filename, mode="exec")
        ns = {**self.kwargs}
        g = self.fn.__globals__
        for k, v in vars(ti).items():
            if k not in g:
                g[k] = v
        exec(code, g, ns)
        if self.fn.__name__ in ns:
            return ns[self.fn.__name__]
        fns = [v for v in ns.values()
               if getattr(v, "__code__", None) is not None and v.__code__.co_filename == self.filename]
        assert len(fns) == 1, list(ns)
        return fns[0]

    ti.FunctionRewriter._compile_and_exec = compile_and_exec
    for name, path in (("sfast", REF), ("sfast.triton", REF + "/triton"),
                       ("sfast.triton.ops", REF + "/triton/ops"), ("sfast.utils", REF + "/utils")):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    gn = importlib.import_module("sfast.triton.ops.group_norm")
    ln = importlib.import_module("sfast.triton.ops.layer_norm")
    assert gn.__file__.startswith(REF) and ln.__file__.startswith(REF)
    return gn, ln


def main():
    gn, ln = _load_reference_kernels()
    out = {"generator": "tests/golden/make_reference_golden.py",
           "reference_files": ["triton/ops/group_norm.py", "triton/ops/layer_norm.py",
                               "triton/ops/activation.py", "triton/ops/utils.py"],
           "group_norm": [], "layer_norm": []}
    g = torch.Generator().manual_seed(20260923)

    def randn(*s):
        return torch.randn(*s, generator=g)

    gn_cases = [
        # (tag, shape, transform of x)
        ("selftest_2x320x32x32", (2, 320, 32, 32), None),
        ("unet_2x1280x8x8", (2, 1280, 8, 8), None),
        ("large_mean_1x320x16x16", (1, 320, 16, 16), "mean50"),
        ("outlier_channel_1x640x8x8", (1, 640, 8, 8), "outlier"),
    ]
    for tag, shape, tf in gn_cases:
        x = randn(*shape)
        if tf == "mean50":
            x = x + 50.0
        elif tf == "outlier":
            x[:, 7] *= 100.0
        x = x.half()
        w, b = randn(shape[1]).half(), randn(shape[1]).half()
        case = {"tag": tag, "x": x, "weight": w, "bias": b, "groups": 32, "eps": 1e-5}
        xcl = x.contiguous(memory_format=torch.channels_last)
        for silu, fn in ((False, gn.group_norm_forward), (True, gn.group_norm_silu_forward)):
            key = "silu" if silu else "plain"
            if tag.startswith("selftest"):
                case["y_nchw_" + key] = fn(x, 32, w, b, 1e-5)[0].contiguous()
            y, mean, rstd = fn(xcl, 32, w, b, 1e-5)
            assert y.is_contiguous(memory_format=torch.channels_last)
            case["y_nhwc_" + key] = y.contiguous()  # stored in NCHW order, values of the NHWC kernel
            case["mean_" + key], case["rstd_" + key] = mean, rstd
        out["group_norm"].append(case)
        print("group_norm", tag, "done", flush=True)

    for tag, (m, n) in (("selftest_recipe_257x1280", (257, 1280)), ("tokens_300x320", (300, 320)),
                        ("tokens_64x640", (64, 640))):
        weight = torch.rand(n, generator=g).half()
        bias = torch.rand(n, generator=g).half()
        x = (-2.3 + 0.5 * randn(m, n)).half()
        y = ln.LayerNorm.forward(types.SimpleNamespace(save_for_backward=lambda *a: None), x, (n,),
                                 weight, bias, 1e-5)
        out["layer_norm"].append({"tag": tag, "x": x, "weight": weight, "bias": bias, "eps": 1e-5, "y": y})
        print("layer_norm", tag, "done", flush=True)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_triton_norms.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
