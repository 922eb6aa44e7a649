#!/usr/bin/env python3
"""Stall samples of one kernel grouped in blocks of N SASS lines, with the marker opcodes seen in each block:
ncu_regions.py file.ncu-rep [kernel-id] [block]"""
import csv, io, subprocess, sys, re
from collections import Counter
path = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "1"
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 250
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, isamp = hdr.index("Source"), hdr.index("# Samples")
stalls = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
half = len(body) // 2 if len(body) > 2 and body[0][isrc] == body[len(body) // 2][isrc] else len(body)
body = body[:half]
tot = sum(int(r[isamp]) for r in body)
print("lines", len(body), "samples", tot)
marks = ["MUFU", "UTCHMMA", "STS", "UBLKCP", "LDTM", "STTM", "LDG", "SYNCS", "UTCBAR", "LDS"]
for b0 in range(0, len(body), blk):
    seg = body[b0:b0 + blk]
    s = sum(int(r[isamp]) for r in seg)
    st = Counter()
    for r in seg:
        for i, h in stalls:
            st[h] += int(r[i] or 0)
    mk = Counter()
    for r in seg:
        for m in marks:
            if re.search(r"\b" + m, r[isrc]):
                mk[m] += 1
    print("%5d-%5d %6d %5.1f%%  %s | %s" % (b0, b0 + len(seg), s, 100.0 * s / tot, " ".join("%s=%d" % kv for kv in st.most_common(4)),
                                            " ".join("%s:%d" % kv for kv in mk.most_common(6))))
