"""Summarise .ncu-rep captures (read here with `ncu -i`, no GPU needed) into one JSON line per kernel:
duration, DRAM bytes / throughput %, tensor-pipe %, issue slots, registers, occupancy, top stall reasons.
Usage: python bench/ncu_summarize.py gpurun_out/ncu_*.ncu-rep > profiles/ncu_summary_r2.jsonl"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_mb",
    "dram__bytes_write.sum": "dram_write_mb",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_throughput_pct",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
}

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        rec = {"report": path.split("/")[-1], "kernel": d.get("Kernel Name", "")[:90]}
        for k, name in KEYS.items():
            if k in d and d[k] != "":
                try:
                    rec[name] = round(float(d[k].replace(",", "")), 3)
                except ValueError:
                    rec[name] = d[k]
        stalls = {h.split("issue_stalled_")[1].split("_per_warp_active")[0]: float(v.replace(",", ""))
                  for h, v in d.items() if "smsp__average_warps_issue_stalled_" in h and h.endswith("_per_warp_active.pct") and v not in ("", "n/a")}
        rec["top_stalls_pct"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        u = dict(zip(hdr, units))
        rec["units"] = {"duration": u.get("gpu__time_duration.sum"), "dram": u.get("dram__bytes_read.sum")}
        print(json.dumps(rec))
