"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum [+ dram__bytes_*]) per kernel family.

usage: python tests/summarize_ncu.py gpurun_out/launches.csv [steps_captured [model-bB-sS [launches_per_step]]] > profiles/rNN_ncu_launch_summary.json
With launches_per_step the first (cold) pass is skipped and the next `steps_captured` complete passes are kept.
The per-launch DRAM traffic of the dominant kernel feeds bench.py's roofline.traffic.
"""
import collections
import csv
import json
import re
import sys


def load(path):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(lines[start:]))


def to_unit(v, unit):
    v = float(v.replace(",", ""))
    scale = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * scale.get(unit, 1.0)


def main():
    rows = load(sys.argv[1])
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    if len(sys.argv) > 4:
        lps = int(sys.argv[4])
        order = sorted({int(r["ID"]) for r in rows})
        keep = set(order[lps:lps * (1 + steps)])
        assert len(keep) == lps * steps, (len(order), lps, steps)
        rows = [r for r in rows if int(r["ID"]) in keep]
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for r in rows:
        fam = re.sub(r"^void ", "", r["Kernel Name"])
        fam = re.sub(r"\(.*", "", fam)
        per[fam][r["Metric Name"]] += to_unit(r["Metric Value"], r["Metric Unit"])
        ids[fam].add(r["ID"])
    out = {"source": sys.argv[1], "steps_captured": steps, "families": {},
           "workload": sys.argv[3] if len(sys.argv) > 3 else "sd15-b2-s64"}
    total_us = sum(m["gpu__time_duration.sum"] for m in per.values())
    for fam, m in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = len(ids[fam])
        d = {"launches": n, "launches_per_step": n / steps,
             "time_us": round(m["gpu__time_duration.sum"], 1),
             "time_share": round(m["gpu__time_duration.sum"] / total_us, 4),
             "avg_us": round(m["gpu__time_duration.sum"] / n, 2)}
        if "dram__bytes_read.sum" in m:
            d["dram_read_bytes_per_launch"] = round(m["dram__bytes_read.sum"] / n)
            d["dram_write_bytes_per_launch"] = round(m["dram__bytes_write.sum"] / n)
            d["dram_bytes_per_launch"] = d["dram_read_bytes_per_launch"] + d["dram_write_bytes_per_launch"]
        out["families"][fam] = d
    out["total_us_per_step"] = round(total_us / steps, 1)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
