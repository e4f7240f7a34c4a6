#!/usr/bin/env python
"""Top stall lines of an `ncu --page source --csv` export: usage ncu_source_top.py file.csv [kernel_index] [n]"""
import csv
import sys


def main(path, which=0, n=25):
    rows = list(csv.reader(open(path)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1] if len(r) > 1 else "?", "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    b = blocks[which]
    hdr = b["hdr"]
    col = {h: i for i, h in enumerate(hdr)}
    g = lambda r, k: int(float(r[col[k]] or 0))
    tot = sum(g(r, "# Samples") for r in b["rows"])
    print(b["name"][:100], "| kernels in file:", len(blocks), "| samples:", tot)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(b["rows"], key=lambda r: -g(r, "# Samples"))[:n]:
        c = g(r, "# Samples")
        st = sorted(((g(r, s), s) for s in stalls), reverse=True)[:3]
        print("%6d %5.1f%%  %-78s %s" % (c, 100.0 * c / max(tot, 1), r[col["Source"]][:78], [(s[6:], v) for v, s in st if v]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
