#!/usr/bin/env python
"""SASS of one kernel of the product library (static view, no GPU needed).
  python scripts/sass_of.py <mangled-name substring> [lib.so]   -> prints the instruction lines; opcode histogram on stderr"""
import collections, re, subprocess, sys
sub = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "opencl-raytracing_b200/libraytracing_cuda.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, keep, hist = None, [], collections.Counter()
for l in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    if cur and sub in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            keep.append(f"{m.group(1)} {m.group(2)}")
            op = re.sub(r"^@!?U?P\d+\s+", "", m.group(2)).split()[0]
            hist[op.split(".")[0]] += 1
print("\n".join(keep))
print(len(keep), "instructions;", ", ".join(f"{k} {v}" for k, v in hist.most_common(14)), file=sys.stderr)
