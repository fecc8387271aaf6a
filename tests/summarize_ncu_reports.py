"""Digest `ncu --set full` reports (gpurun_out/r02_prof_*.ncu-rep) into one JSON of the counters the
roofline discussion uses:  python tests/summarize_ncu_reports.py gpurun_out/r02_prof_*.ncu-rep > profiles/rNN_ncu_kernel_details.json"""
import csv
import io
import json
import os
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active": "tensor_issue_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_active_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_mufu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__warps_active.avg.per_cycle_active": "warps_active_per_smsp",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_bytes",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio": "stall_membar",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
}
SCALE = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def digest(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return [{"error": "no kernels in report"}]
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")][:120]}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT and v not in ("", "n/a"):
                try:
                    d[WANT[h]] = round(float(v.replace(",", "")) * SCALE.get(u, 1.0), 3)
                except ValueError:
                    pass
        res.append(d)
    return res


if __name__ == "__main__":
    json.dump({os.path.basename(p).replace(".ncu-rep", ""): digest(p) for p in sys.argv[1:]}, sys.stdout, indent=1)
