#!/usr/bin/env python3
"""tools/ncu_lines.py REPORT.ncu-rep|LINES.csv [top] -- warp-instructions executed and stall samples per CUDA source line of the
(first) kernel in an `ncu --set full --import-source on` report.  Runs here, no GPU needed."""
import csv, io, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = (open(rep, encoding="latin-1").read() if rep.endswith(".csv") else
       subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True).stdout.decode("latin-1"))
fname, hdr, lines = "", None, []
for r in csv.reader(io.StringIO(raw)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        try:
            lines.append((int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")] or 0), fname, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot = sum(l[0] for l in lines) or 1
smp = sum(l[1] for l in lines) or 1
print("%.3e warp-instructions, %d samples" % (tot, smp))
for v, s, f, ln, src in sorted(lines, key=lambda x: -x[0])[:top]:
    print("%5.1f%% smp %5.1f%%  %s:%d  %s" % (100.0 * v / tot, 100.0 * s / smp, f, ln, src[:100]))
