#!/usr/bin/env python
"""profiles/ncu_counters.json from an `ncu --set full` capture: per hot kernel, the counters bench.py quotes next to its
roofline blocks (issue-slot fraction and DRAM traffic beside the algorithmic GB/s).

  ncu --set full --clock-control none -k regex:'k_(shadow|extend|shade)' -c 27 -o gpurun_out/r3_c3 python scripts/profile_target.py ...
  python scripts/ncu_counters.py C3 gpurun_out/r3_c3.ncu-rep [profiles/ncu_counters.json]

Per kernel (all captured launches of it, duration-weighted): issue slots busy %, lanes per warp instruction, DRAM bytes per
launch (read + write) and % of peak, L2 hit rate, the mean launch duration under ncu (cold caches, serialised: compare the
kernel's share with bench.py's CUDA-event figure, not the absolute), and `bound`: "hbm" when DRAM throughput is the busiest
unit, else "issue".
"""
import csv
import json
import os
import subprocess
import sys

M = {"dur": "gpu__time_duration.sum", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l2hit": "lts__t_sector_hit_rate.pct",
     "l1hit": "l1tex__t_sector_hit_rate.pct", "occ": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "regs": "launch__registers_per_thread", "warp_inst": "smsp__inst_executed.sum"}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def main():
    workload, rep = sys.argv[1], sys.argv[2]
    # the capture runs the workload's frame at a reduced spp (one batch either way): every launch then carries spp_bench / spp_profiled
    # times fewer items, and the per-launch figures bench.py compares with (DRAM bytes, duration) are scaled by that ratio
    scale = 1.0
    for a in list(sys.argv[3:]):
        if a.startswith("--scale="):
            scale = float(a.split("=")[1])
            sys.argv.remove(a)
    out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_counters.json")
    # either an .ncu-rep, or the text of `ncu -i <rep> --page raw --csv` made on the GPU box (reports of many launches exceed what
    # a GPU session may bring back)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        i = idx.get(M[key])
        if i is None or r[i] in ("", "n/a"):
            return None
        v = float(r[i].replace(",", ""))
        return v * UNIT_SCALE.get(units[i], 1.0)

    per = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        short = "k_shadow_gather" if "k_shadow_gather" in name else "k_shadow" if "k_shadow" in name else "k_extend" if "k_extend" in name else \
            "k_shade" if "k_shade" in name else None
        if short is None or short == "k_shadow_gather":
            continue
        per.setdefault(short, []).append(r)
    result = {}
    for k, rs in per.items():
        w = [val(r, "dur") or 0.0 for r in rs]
        tot = sum(w) or 1.0

        def avg(key):
            vals = [(val(r, key), wi) for r, wi in zip(rs, w) if val(r, key) is not None]
            return sum(v * wi for v, wi in vals) / (sum(wi for _, wi in vals) or 1.0) if vals else None
        dram = sum((val(r, "rd") or 0.0) + (val(r, "wr") or 0.0) for r in rs) / len(rs)
        issue, dram_pct = avg("issue"), avg("dram_pct")
        result[k] = {"launches_captured": len(rs), "ncu_ms_per_launch": scale * tot / len(rs), "issue_active_pct": issue, "lanes_per_inst": avg("lanes"),
                     "dram_bytes_per_launch": scale * dram, "per_launch_scale": scale, "dram_pct_of_peak": dram_pct, "l2_hit_pct": avg("l2hit"), "l1_hit_pct": avg("l1hit"),
                     "achieved_occupancy_pct": avg("occ"), "registers": avg("regs"), "warp_instructions_per_launch": sum(val(r, "warp_inst") or 0 for r in rs) / len(rs),
                     "bound": "hbm" if (dram_pct or 0) > (issue or 0) else "issue", "ncu_source": os.path.basename(rep)}
    allc = json.load(open(out_path)) if os.path.exists(out_path) else {}
    allc[workload] = result
    json.dump(allc, open(out_path, "w"), indent=1, sort_keys=True)
    print(json.dumps(result, indent=1))


if __name__ == "__main__":
    main()
