#!/usr/bin/env python
"""Benchmark of the UNet denoise hot path (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # B200-native arm
    python bench.py --impl reference --gpus N ...            # reference CPU path (oracle port)

One "step" = one UNet forward over this rank's batch of synthetic latents (default: the CFG
pair, B = 2, of SD-1.5 at 512x512 -> 4x64x64 latents, fp16, CUDA graph on = BASELINE.json
configs[1]).  `value` = latents pushed through one UNet forward per second, whole job
(N GPUs x B x K / max-over-ranks device time).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "stable-fast_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "unet_latent_steps_per_s"
UNIT = "latents through one UNet forward per second (SD-1.5 20-step ms/img = 2*20*1000/value)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="sd15", choices=["sd15", "sdxl", "tiny"])
    ap.add_argument("--batch", type=int, default=2, help="latents per GPU per step")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: total latents sharded over the GPUs (BASELINE configs[4])")
    ap.add_argument("--size", type=int, default=64, help="latent height = width")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--dump-ops", default="", help="write per-op CUDA-event timings (JSON lines)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons,
                "samples": len(rows)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured"
    return 1400.0, 6650.0, "fallback"


def usable_cores():
    """Host threads this process may really use: affinity mask, capped by a cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                quota, period = txt[0], float(txt[1])
            else:
                quota, period = txt[0], float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota not in ("max", "-1"):
                n = max(1, min(n, int(float(quota) / period)))
            break
        except (OSError, ValueError, IndexError):
            continue
    return n


def workload_name(args, batch, ctx_dim):
    return (f"{args.model} UNet2DConditionModel forward, {batch} latent(s) 4x{args.size}x{args.size} "
            f"per step per GPU (B=2 = the CFG pair of one {args.size * 8}^2 image), text 77x{ctx_dim}, "
            "random-init weights")


def cross_dim(args):
    from sfast_b200.synthetic import CONFIGS
    return CONFIGS[args.model]["cross_attention_dim"]


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(args, steps, warmup, budget_s):
    """The reference's CPU path for this workload = plain eager PyTorch fp32 (every sfast op falls
    back to aten on CPU: /root/reference/src/sfast/csrc/operators/cublas/cublas_gemm.cpp:705-710,
    /root/reference/src/sfast/triton/torch_ops.py:116-126), run on the oracle restatement of the
    UNet because diffusers is not installable here.  One step = the same batch as the GPU arm."""
    from oracle import unet_oracle as uo
    cores = usable_cores()
    torch.set_num_threads(cores)
    cfg = {"sd15": uo.sd15_config, "sdxl": uo.sdxl_config, "tiny": uo.tiny_config}[args.model]()
    m = uo.build_unet(cfg, seed=0)
    g = torch.Generator().manual_seed(0)
    nb = args.batch
    s = torch.randn(nb, 4, args.size, args.size, generator=g)
    e = torch.randn(nb, 77, cfg.cross_attention_dim, generator=g)
    kw = {}
    if cfg.addition_embed_type == "text_time":
        kw["added_cond_kwargs"] = {"text_embeds": torch.randn(nb, 1280, generator=g),
                                   "time_ids": torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * nb)}
    times = []
    with torch.no_grad():
        t0 = time.perf_counter()
        for _ in range(max(1, min(warmup, 1))):
            m(s, torch.tensor(999.0), e, **kw)
        t_warm = time.perf_counter() - t0
        n = max(1, min(steps, int(budget_s / max(t_warm, 1e-3))))
        for i in range(n):
            t0 = time.perf_counter()
            m(s, torch.tensor(999.0 - 50 * i), e, **kw)
            times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"value": nb / med, "ms_per_step": med * 1e3, "steps": n, "cores": cores,
            "sample": f"{n} timed fp32 eager forward(s) of {nb} {args.model} latent(s) "
                      f"4x{args.size}x{args.size} (the same batch as the GPU arm) after 1 warm-up, "
                      f"{cores} threads, median"}


def run_reference(args, rank):
    if rank != 0:
        return
    r = cpu_reference_run(args, args.steps, args.warmup, budget_s=100.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": 1, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args, args.batch, cross_dim(args)),
                   "global_batch": args.batch, "parallelism": "cpu",
                   "arm": "fp32 CPU eager (the reference's CPU path = aten fallbacks) on the oracle port"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local):
    from sfast.compilers.diffusion_pipeline_compiler import CompilationConfig, compile_unet
    from sfast_b200 import _lib, dist as sdist
    from sfast_b200.synthetic import CONFIGS, SyntheticUNet
    from sfast_b200.unet_spec import param_shapes, random_state_dict, spec_from_config

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dtype = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    cfg = CONFIGS[args.model]
    spec = spec_from_config(cfg)
    if args.global_batch:
        lo, hi = sdist.shard_batch(args.global_batch, rank, world)
        batch, scaling = hi - lo, "strong"
    else:
        batch, scaling = args.batch, "weak"
    # weights: rank 0 draws them, ONE broadcast over NCCL/NVLink, zero collectives afterwards
    sd0 = random_state_dict(spec, seed=0, dtype=dtype, device=dev) if rank == 0 else None
    sd = sdist.broadcast_state_dict(param_shapes(spec), sd0, dtype, dev) if world > 1 else sd0
    unet = SyntheticUNet(cfg, state_dict=sd, dtype=dtype, device=dev)
    cc = CompilationConfig.Default()
    cc.enable_cuda_graph = not args.no_graph
    unet = compile_unet(unet, cc)
    compiled = unet.forward._compiled

    ctx_dim = spec.cross_attention_dim
    g = torch.Generator().manual_seed(1234 + rank)
    n_var = 4  # a few distinct host inputs so steps are not identical
    host_s = [torch.randn(batch, 4, args.size, args.size, generator=g).to(dtype).pin_memory()
              for _ in range(n_var)]
    host_e = [torch.randn(batch, 77, ctx_dim, generator=g).to(dtype).pin_memory()
              for _ in range(n_var)]
    host_out = torch.empty(batch, 4, args.size, args.size, dtype=dtype).pin_memory()
    kw = {}
    if spec.addition_embed_type == "text_time":
        kw["added_cond_kwargs"] = {
            "text_embeds": torch.randn(batch, 1280, generator=g).to(dev, dtype),
            "time_ids": torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * batch, device=dev)}
    dev_s = [t.to(dev) for t in host_s]
    dev_e = [t.to(dev) for t in host_e]
    tsteps = [torch.tensor(float(999 - 50 * (i % 20)), device=dev) for i in range(20)]

    lib = _lib.lib()
    n0 = lib.sfb_launch_count()
    unet(dev_s[0], tsteps[0], dev_e[0], **kw)  # builds plan (+ graph): eager warm-up + capture
    torch.cuda.synchronize()
    gp = next(iter(compiled._cached.values()))
    plan = gp.plan
    # first call = eager warm-up pass + (graph capture | eager step): two passes over the plan
    launches_per_step = (lib.sfb_launch_count() - n0) // 2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = sdist.max_over_ranks(e0.elapsed_time(e1), dev)
        return ms, clocks

    # --- device-resident arm: inputs already in HBM; per step = static-buffer copies + graph replay
    def step_dev(i):
        unet(dev_s[i % n_var], tsteps[i % 20], dev_e[i % n_var], **kw)

    ms_dev, clocks = timed(step_dev, args.steps, max(args.warmup, 3))

    # --- end-to-end arm: pinned host inputs -> H2D -> public API call -> D2H of the result
    def step_e2e(i):
        s = host_s[i % n_var].to(dev, non_blocking=True)
        e = host_e[i % n_var].to(dev, non_blocking=True)
        out = unet(s, tsteps[i % 20], e, **kw).sample
        host_out.copy_(out, non_blocking=True)

    ms_e2e, _ = timed(step_e2e, args.steps, max(args.warmup, 3))

    weight_gb = sum(int(torch.Size(sh).numel()) for sh in param_shapes(spec).values()) * 2 / 1e9
    total_latents = batch * world if not args.global_batch else args.global_batch
    value = total_latents * args.steps / (ms_dev / 1e3)
    e2e_value = total_latents * args.steps / (ms_e2e / 1e3)
    h2d = host_s[0].numel() * 2 + host_e[0].numel() * 2
    d2h = host_out.numel() * 2

    roofline = None
    if not args.no_roofline:
        roofline = measure_roofline(plan, lib, args.dump_ops, f"{args.model}-b{batch}-s{args.size}")
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args, steps=2, warmup=1, budget_s=25.0)
        cpu_base = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                    "sample": r["sample"]}
    if rank != 0:
        return
    ms_step = ms_dev / args.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {
            "workload": workload_name(args, batch, ctx_dim),
            "cuda_graph": not args.no_graph,
            "global_batch": total_latents, "parallelism": f"dp{world}",
            "l2": f"inputs larger than L2: the {weight_gb:.2f} GB 16-bit weight set streams from HBM every step",
            "sd15_20step_unet_ms_per_img": 20 * ms_step if (args.model == "sd15" and batch == 2) else None,
            "algorithmic_tflop_per_step": plan.flops() / 1e12,
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches_per_step * args.steps),
        "kernel_launches_per_step": int(launches_per_step),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_base,
    }
    print(json.dumps(line), flush=True)


def ncu_traffic(prefix, workload=None):
    """DRAM bytes per launch of the kernel family, from the committed ncu launch list of this same
    command (profiles/rNN_ncu_launch_summary.json, made by tests/summarize_ncu.py; ncu flushes the
    caches before every kernel, so this is cold-cache traffic)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_launch_summary.json")))
    if not files:
        return {"traffic": None}
    d = json.load(open(files[-1]))
    if workload is not None and d.get("workload") != workload:
        return {"traffic": None, "traffic_note": f"no ncu launch list committed for {workload}"}
    n = b = 0
    for name, f in d.get("families", {}).items():
        if name.startswith(prefix) and "dram_bytes_per_launch" in f:
            n += f["launches"]
            b += f["dram_bytes_per_launch"] * f["launches"]
    if not n:
        return {"traffic": None}
    return {"traffic": b / n, "traffic_unit": "DRAM bytes per launch (read + write, ncu, cold caches)",
            "traffic_source": os.path.relpath(files[-1], ROOT)}


def measure_roofline(plan, lib, dump_path="", workload=None):
    """Per-op CUDA-event timing of one eager pass over the plan (after the timed region).  The
    dominant kernel is the tcgen05 GEMM / implicit-GEMM conv; its roofline is the tensor pipe."""
    stream = torch.cuda.current_stream()
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
    evs = []
    for op in plan.all_ops():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        op.launch(stream.cuda_stream)
        b.record(stream)
        evs.append((op, a, b))
    torch.cuda.synchronize()
    if dump_path:
        with open(dump_path, "w") as f:
            for op, a, b in evs:
                extra = {}
                if getattr(op.fn, "__name__", "") == "sfb_gemm":
                    p = op.keep[0]
                    extra = {"M": p.M, "N": p.N, "K": p.K, "splits": p.splits, "conv": p.a_mode,
                             "epi": p.epi}
                f.write(json.dumps({"op": op.name, "fn": getattr(op.fn, "__name__", "?"),
                                    "us": a.elapsed_time(b) * 1e3, "flops": op.flops,
                                    "bytes": op.bytes, **extra}) + "\n")
    # steady-state cost of the dominant kernel family: all of the step's sfb_gemm launches (same
    # buffers, same order) captured in a CUDA graph and replayed, timed with CUDA events
    gemm_ops = [op for op, _, _ in evs if getattr(op.fn, "__name__", "") == "sfb_gemm"]
    gg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gg):
        cs = torch.cuda.current_stream().cuda_stream
        for op in gemm_ops:
            op.launch(cs)
    for _ in range(3):
        gg.replay()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(10):
        gg.replay()
    g1.record()
    torch.cuda.synchronize()
    gemm_graph_ms = g0.elapsed_time(g1) / 10
    fam = {}
    for op, a, b in evs:
        key = getattr(op.fn, "__name__", None) or str(op.name)
        d = fam.setdefault(key, {"ms": 0.0, "flops": 0, "bytes": 0, "n": 0})
        d["ms"] += a.elapsed_time(b)
        d["flops"] += op.flops
        d["bytes"] += op.bytes
        d["n"] += 1
    total_ms = sum(d["ms"] for d in fam.values())
    peak_tf, peak_bw, src = load_peaks()
    g = fam.get("sfb_gemm", {"ms": 1e-9, "flops": 0, "n": 1, "bytes": 0})
    achieved = g["flops"] / (gemm_graph_ms / 1e3) / 1e12
    shares = {k: round(v["ms"] / total_ms, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    att = fam.get("sfb_attention")
    extra = {}
    if att:
        extra["attention_tflops"] = att["flops"] / (att["ms"] / 1e3) / 1e12
    gn = fam.get("sfb_group_norm_apply")
    if gn:
        extra["group_norm_apply_gbs"] = gn["bytes"] / (gn["ms"] / 1e3) / 1e9
    return {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 GEMM / implicit-GEMM 3x3 conv, all instances)",
            "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
            "peak_source": f"{src} (MEASURED_PEAKS.json bf16_tflops_sustained)",
            "launches_per_step": g["n"], "avg_launch_us": gemm_graph_ms * 1e3 / max(g["n"], 1),
            "timing": "all sfb_gemm launches of one step replayed as a CUDA graph, CUDA events, 10 replays",
            "avg_launch_us_eager_events": g["ms"] * 1e3 / max(g["n"], 1),
            "flop_per_step": g["flops"], **ncu_traffic("gemm_tc_kernel", workload),
            "eager_step_ms": total_ms, "time_share_by_entry_point": shares, **extra}


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the B200 arm needs a CUDA device (no CPU fallback); "
                         "use --impl reference for the CPU baseline")
    if world > 1:
        from sfast_b200 import dist as sdist
        sdist.init_from_env("nccl")
    run_b200(args, rank, world, local)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
