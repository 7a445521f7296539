#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an ncu report: joins `ncu --page source --csv` (SASS rows with stall samples
and executed thread instructions) with `nvdisasm --print-line-info` of the cubin inside the given object / .so.

  python scripts/ncu_hotspots.py <report.ncu-rep> <kernel regex> <object with the cubin> [top N]
"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, kre, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre.split("<")[0]], capture_output=True, text=True).stdout
lines = raw.splitlines()
# first kernel block whose full name contains the requested text (k_shadow< vs k_shadow_gather)
nth = int(os.environ.get("NTH", "0"))   # which launch of the kernel in the report
start = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"') and kre in l][nth]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
kname = next(csv.reader([lines[start]]))[1]
rows = list(csv.reader(lines[start + 1:end]))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
sass = rows[1:]
base = int(sass[0][ix["Address"]], 16)

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
mangled = None
# find the section whose demangled name matches: use the mangled name fragment from the kernel name
frag = re.sub(r"^(void )?(rt::)?", "", kname).split("<")[0].split("(")[0]
loc, cur, infn = {}, ("?", 0), False
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        sec = m.group(1)   # mangled names carry the identifier length: 8k_shadow vs 15k_shadow_gather
        infn = f"{len(frag)}{frag}" in sec and ("ILb1" not in sec) and (("DiffuseSurface" in sec) == ("DiffuseSurface" in kname))
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m: loc[int(m.group(1), 16)] = cur

agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in sass:
    off = int(r[ix["Address"]], 16) - base
    key = loc.get(off, ("?", 0))
    vals = [int(float(r[ix[k]] or 0)) for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed")]
    for j in range(3):
        agg[key][j] += vals[j]; tot[j] += vals[j]
print(f"kernel {kname}: {tot[0]} samples, {tot[1]} warp inst, {tot[2]} thread inst, {tot[2] / max(1, tot[1]):.1f} lanes/inst")
print(f"{'file:line':32s} {'samples%':>8s} {'warp inst%':>10s} {'lanes':>6s}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0] + ':' + str(key[1]):32s} {100 * v[0] / max(1, tot[0]):8.2f} {100 * v[1] / max(1, tot[1]):10.2f} {v[2] / max(1, v[1]):6.1f}")
# per file rollup
f = collections.defaultdict(lambda: [0, 0, 0])
for key, v in agg.items():
    for j in range(3): f[key[0]][j] += v[j]
print("-- per file")
for k, v in sorted(f.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:32s} {100 * v[0] / max(1, tot[0]):8.2f} {100 * v[1] / max(1, tot[1]):10.2f} {v[2] / max(1, v[1]):6.1f}")
