"""CPU tests of bench.py's contract pieces that need no GPU: the reference arm's JSON line (run as the
driver runs it, on the tiny config) and the launch-list lookup behind `roofline.traffic`."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny",
                          "--size", "16", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["higher_is_better"] is True and d["value"] > 0 and d["data"] == "synthetic"
    for key in ("metric", "unit", "ms_per_step", "dtype", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 == d["e2e"]["d2h_bytes_per_step"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_ncu_traffic_picks_the_launch_list_of_its_workload():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        sys.argv = argv
    a = m.ncu_traffic("gemm_tc_kernel", "sd15-b2-s64")
    b = m.ncu_traffic("gemm_tc_kernel", "sd15-b8-s64")
    assert a["traffic"] and a["traffic_source"].endswith("_ncu_launch_summary.json")
    assert b["traffic"] and b["traffic_source"].endswith("_ncu_launch_summary_b8.json") and b["traffic"] > a["traffic"]
    none = m.ncu_traffic("gemm_tc_kernel", "no-such-workload")
    assert none["traffic"] is None and "no-such-workload" in none["traffic_note"]
