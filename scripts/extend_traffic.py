#!/usr/bin/env python
"""profiles/extend_traffic.json from an ncu CSV of the bench's k_extend launches (roofline.traffic in bench.py):

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \\
      -k regex:k_extend -c 72 --csv --log-file gpurun_out/extend_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline
  python scripts/extend_traffic.py gpurun_out/extend_traffic.csv "<note>"

72 launches = the 8 wavefront batches x 9 depths of one C3 frame; the value is the mean DRAM bytes (read + write) per launch."""
import csv, json, os, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(r for r in rows if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows:
    if len(r) != len(hdr) or r is hdr or not r[ix["ID"]].isdigit():
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(unit, 1)
    per.setdefault(int(r[ix["ID"]]), {})[r[ix["Metric Name"]]] = v * scale
n = len(per)
rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values())
wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values())
t = sum(p.get("gpu__time_duration.sum", 0) for p in per.values())
out = {"kernel": "k_extend", "launches": n, "dram_bytes_per_launch": (rd + wr) / max(1, n), "dram_read_bytes_per_launch": rd / max(1, n),
       "dram_write_bytes_per_launch": wr / max(1, n), "ms_per_launch_under_ncu": 1e3 * t / max(1, n),
       "source": os.path.basename(sys.argv[1]), "note": sys.argv[2] if len(sys.argv) > 2 else ""}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, "profiles", "extend_traffic.json"), "w"), indent=1)
print(json.dumps(out))
