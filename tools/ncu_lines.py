#!/usr/bin/env python
"""Stall samples per CUDA source line of one launch in an .ncu-rep (captured with --import-source on, built with -lineinfo).
   tools/ncu_lines.py rep.ncu-rep <launch index> [min_pct]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2])
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(k),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, hdr, out, tot = None, None, [], 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": print(r[1][:150]); continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; i_s = hdr.index("# Samples"); i_ex = hdr.index("Instructions Executed"); continue
    if hdr and len(r) > i_s and r[0].isdigit() and r[i_s].isdigit():
        n = int(r[i_s]); tot += n
        out.append((cur, int(r[0]), n, int(r[i_ex]) if r[i_ex].isdigit() else 0, r[1][:130]))
print("total samples", tot)
for f, l, n, ex, s in out:
    if n >= tot * min_pct / 100: print(f"{f}:{l:4d} {100*n/tot:5.1f}% ex={ex:9d} {s}")
