#!/usr/bin/env python
"""Per-instruction shared-memory view of an ncu report (--import-source on): which instructions make the shared-memory
wavefronts and the "bank conflict" count.  Prints every column of the source page whose name mentions shared memory or
wavefronts, for the instructions that have a non-zero value in one of them.
    python profiles/shared_access.py gpurun_out/prof.ncu-rep [kernel-index]"""
import csv, io, subprocess, sys

path = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None and row:
        cur["rows"].append(row)
b = blocks[kidx]
hdr = b["rows"][0]
rows = [r for r in b["rows"][1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
cols = [h for h in hdr if any(w in h.lower() for w in ("shared", "wavefront", "conflict"))]
print("kernel:", b["name"])
print("columns:", cols)
def num(v):
    try: return float(v)
    except ValueError: return 0.0
tot = {c: 0.0 for c in cols}
hits = []
for r in rows:
    vals = [num(r[ix[c]]) for c in cols]
    if any(vals):
        hits.append((r, vals))
        for c, v in zip(cols, vals): tot[c] += v
hits.sort(key=lambda h: -max(h[1]))
for r, vals in hits[:40]:
    print("  exec=%-10s %-58s %s" % (r[ix["Instructions Executed"]], r[ix["Source"]][:58], "  ".join("%s=%d" % (c[:34], v) for c, v in zip(cols, vals) if v)))
print("totals:", {c: int(v) for c, v in tot.items()})
