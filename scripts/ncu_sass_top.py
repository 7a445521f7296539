#!/usr/bin/env python
"""Top SASS instructions of one kernel in an ncu report by stall samples, with the dominant stall reasons and the
instructions around them (no cubin needed).   python scripts/ncu_sass_top.py <report.ncu-rep> <kernel substring> [top N] [context]"""
import csv, os, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
nth = int(os.environ.get("NTH", "0"))   # which launch of the kernel in the report
start = [i for i, l in enumerate(raw) if l.startswith('"Kernel Name"') and kre in l][nth]
end = next((i for i in range(start + 1, len(raw)) if raw[i].startswith('"Kernel Name"')), len(raw))
rows = list(csv.reader(raw[start + 1:end]))
hdr, sass = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def num(r, k):
    try: return float(r[ix[k]] or 0)
    except ValueError: return 0.0
tot = sum(num(r, "# Samples") for r in sass)
tot_inst = sum(num(r, "Instructions Executed") for r in sass)
print(f"{next(csv.reader([raw[start]]))[1]}: {len(sass)} SASS instructions, {tot:.0f} samples, {tot_inst:.3g} warp instructions")
agg = {s: sum(num(r, s) for r in sass) for s in stalls}
print("stall totals:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot))
order = sorted(range(len(sass)), key=lambda i: -num(sass[i], "# Samples"))[:top]
for i in order:
    r = sass[i]
    why = sorted(((num(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{100 * num(r, '# Samples') / tot:5.2f}%  #{i:5d} lanes {num(r, 'Avg. Threads Executed'):4.1f} exec {100 * num(r, 'Instructions Executed') / tot_inst:4.2f}%  "
          f"{why[0][1]} {why[0][0]:.0f}, {why[1][1]} {why[1][0]:.0f}   {r[ix['Source']][:90]}")
    for j in range(max(0, i - ctx), i):
        print(f"          #{j:5d} {sass[j][ix['Source']][:100]}")
