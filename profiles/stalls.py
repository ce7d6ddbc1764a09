#!/usr/bin/env python
"""Per-instruction view of an ncu report: top stall sites and stall-reason totals for one kernel.
    python profiles/stalls.py gpurun_out/prof.ncu-rep [kernel-index] [top-n]"""
import csv, io, subprocess, sys, collections

path = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
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
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in rows:
    for c in stall_cols:
        try: tot[c] += float(r[ix[c]])
        except ValueError: pass
allsamp = sum(tot.values())
print("kernel:", b["name"])
print("instructions:", len(rows), " warp-level executed:", sum(float(r[ix["Instructions Executed"]] or 0) for r in rows))
print("stall reasons (all samples):")
for c, v in tot.most_common(12):
    print("   %-28s %8.0f  %5.1f%%" % (c, v, 100 * v / max(allsamp, 1)))
print("top stall sites:")
rows.sort(key=lambda r: -float(r[ix["# Samples"]] or 0))
for r in rows[:topn]:
    reasons = sorted(((float(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print("   %6s samples  exec=%-9s %-60s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]][:60],
                                                  ", ".join("%s=%d" % (c[6:], v) for v, c in reasons if v)))
