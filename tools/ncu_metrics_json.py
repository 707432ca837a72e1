"""profiles/bcp_kernel_metrics.json from the round's .ncu-rep files (read here, no GPU needed).
usage: python tools/ncu_metrics_json.py key=rep:n_samples ...   (first bcp_kernel launch of each report)"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {
    "gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_instruction",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflict_wavefronts",
    "sm__cycles_elapsed.max": "sm_cycles", "launch__registers_per_thread": "registers_per_thread",
    "launch__block_size": "block_size", "launch__grid_size": "grid_size",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}

out = {}
for spec in sys.argv[1:]:
    key, rest = spec.split("=")
    rep, n = rest.split(":")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    row = [r for r in rows[2:] if "bcp_kernel" in r[hdr.index("Kernel Name")]][0]
    d = {"samples": int(n), "source": os.path.basename(rep), "kernel": row[hdr.index("Kernel Name")].split("(")[0]}
    for k, name in KEYS.items():
        if k in hdr:
            i = hdr.index(k)
            v = float(row[i].replace(",", ""))
            v *= UNIT_SCALE.get(units[i], 1)
            d[name] = v
    d["dram_bytes_per_sample"] = (d.get("dram_bytes_read", 0) + d.get("dram_bytes_write", 0)) / int(n)
    d["dram_gbs"] = (d.get("dram_bytes_read", 0) + d.get("dram_bytes_write", 0)) / (d["time_us"] * 1e-6) / 1e9
    d["warp_instructions_per_group"] = d["warp_instructions"] / ((int(n) + 63) // 64)
    d["smem_wavefronts_per_cycle_per_sm"] = d["smem_wavefronts"] / (d["sm_cycles"] * d["grid_size"]) if d.get("grid_size") else None
    out[key] = d
json.dump(out, open(os.path.join(ROOT, "profiles", "bcp_kernel_metrics.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
