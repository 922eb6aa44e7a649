#!/usr/bin/env python3
"""Top stall-sample SASS lines of one kernel of an ncu report: ncu_hot.py file.ncu-rep [kernel-id] [N]"""
import csv, io, subprocess, sys
path = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "1"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for ln, r in enumerate(rows[2:]):
    if len(r) <= isamp:
        continue
    try:
        s = int(r[isamp])
    except ValueError:
        continue
    data.append((s, ln, r))
tot = sum(d[0] for d in data)
print("total samples", tot)
for s, ln, r in sorted(data, reverse=True)[:N]:
    top = sorted(((int(r[i] or 0), h) for i, h in stalls), reverse=True)[:2]
    print("%6d %5.1f%% line %5d  %-60s %s" % (s, 100.0 * s / tot, ln, r[isrc].strip()[:60], " ".join("%s=%d" % (h[6:], v) for v, h in top if v)))
