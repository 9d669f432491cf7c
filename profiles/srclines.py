#!/usr/bin/env python
"""Per-source-line warp-stall samples of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
   python profiles/srclines.py gpurun_out/prof.ncu-rep regex:wb_omega [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, agg, hdr = None, [], None
tot_reason = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr and len(r) > 6 and r[2] == "-":
        try:
            ns = int(r[hdr.index("# Samples")])
            reasons = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:3]
            for i, h in stall_cols:
                tot_reason[h[6:]] = tot_reason.get(h[6:], 0) + int(r[i] or 0)
            agg.append((ns, int(r[7]), fname, r[0], r[1].strip()[:90], reasons))
        except ValueError:
            pass
tot = sum(a[0] for a in agg)
print(f"# {kern}: {tot} samples; stall reasons overall: " +
      ", ".join(f"{k} {v / max(tot, 1):.2f}" for k, v in sorted(tot_reason.items(), key=lambda kv: -kv[1])[:8]))
print(f"# top {top} source lines (samples, share, warp-instructions, file:line, top stall reasons, source)")
for s, n, f, ln, src, reasons in sorted(agg, reverse=True)[:top]:
    rs = " ".join(f"{h}:{v}" for v, h in reasons if v)
    print(f"{s:7d} {s / max(tot, 1):6.3f} {n:10d}  {f}:{ln:>4s}  [{rs}]  {src}")
