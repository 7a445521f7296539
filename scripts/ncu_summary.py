#!/usr/bin/env python
"""Summarise an ncu report (read with `ncu -i <rep> --page raw --csv`) into a small markdown table for profiles/."""
import csv, subprocess, sys
KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thread"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / warp instruction"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles / issued inst"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
        ("smsp__inst_executed_op_local_ld.sum", "local loads"), ("smsp__inst_executed_op_local_st.sum", "local stores")]
rep = sys.argv[1]
# either an .ncu-rep, or the text of `ncu -i <rep> --page raw --csv` made on the GPU box; optional: the launch IDs to tabulate
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
kern = [r for r in rows[2:] if want is None or r[idx["ID"]] in want]
print("| metric | " + " | ".join(r[idx["Kernel Name"]].split("(")[0].replace("void ", "") for r in kern) + " |")
print("|---|" + "---|" * len(kern))
for k, label in KEYS:
    if k not in idx: continue
    def fmt(r):
        v = r[idx[k]]
        try: v = f"{float(v.replace(',', '')):.4g}"
        except ValueError: pass
        return f"{v} {units[idx[k]]}".strip()
    print(f"| {label} (`{k}`) | " + " | ".join(fmt(r) for r in kern) + " |")
