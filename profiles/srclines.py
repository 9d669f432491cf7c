#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
   python profiles/srclines.py gpurun_out/prof.ncu-rep regex:wb_omega [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, agg, hdr, seen_fn = None, [], None, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        seen_fn += 1
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) > 6 and r[2] == "-" and seen_fn <= 99:
        try:
            agg.append((int(r[hdr.index("# Samples")]), int(r[7]), fname, r[0], r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(a[0] for a in agg)
print(f"# {kern}: {tot} samples; top {top} source lines (samples, share, warp-instructions, file:line, source)")
for s, n, f, ln, src in sorted(agg, reverse=True)[:top]:
    print(f"{s:7d} {s / max(tot, 1):6.3f} {n:10d}  {f}:{ln:>4s}  {src}")
