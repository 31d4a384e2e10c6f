#!/usr/bin/env python3
"""tools/path_bench.py [GiB] -- development timing of the scan path for a spread of patterns (count-only, device-resident
synthetic corpus): which plan they get, stage times, GB/s.  Not part of the driver contract (that is bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import agrep_b200 as ag

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gib * (1 << 30)) // 4096 * 4096
t = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
t[n:].zero_()
ag.corpus_device(t.data_ptr(), n, needle="because each", needle_every=4096, needle_maxedits=3, paragraphs=("para" in sys.argv))
torch.cuda.synchronize()
CASES = [
    ("the", dict()), ("government", dict()), ("because each", dict(k=0, linenum=True)), ("because each", dict(k=1)),
    ("because each", dict(k=2)), ("because each", dict(k=3)), ("because each", dict(k=2, nocase=True)),
    ("because each just those", dict(k=4, nocase=True, linenum=True)),
    ("business give group toward young", dict(k=3, wordbound=True, linenum=True)),
    ("because each", dict(k=2, inverse=True)), ("b[ea]cause", dict(k=0, linenum=True)), ("b[ea]c.u[s-t]e", dict(k=1, linenum=True)),
    ("th", dict(k=1, linenum=True)), ("because each", dict(k=2, cost_s=2)), ("governmental", dict(k=8, linenum=True)),
]
print("%-36s %-22s plan NA refine | front ms  stage2 ms   total GB/s   matched    flagged%%" % ("pattern", "options"))
for pat, kw in CASES:
    p = ag.Pattern(pat, **kw)
    d = p.desc
    p.scan_device(t.data_ptr(), n)
    best = None
    for _ in range(3):
        r = p.scan_device(t.data_ptr(), n)
        tot = r.ms_front + r.ms_records
        if best is None or tot < best[0]:
            best = (tot, r.ms_front, r.ms_records, r.n_matched, r.n_flagged)
    print("%-36r %-22s %4d %2d %6d | %8.3f %10.3f %10.1f %10d %9.3f" % (
        pat, ",".join("%s=%s" % (k, int(v)) for k, v in kw.items()), d.plan, d.n_anchors, d.refine,
        best[1], best[2], n / best[0] / 1e6, best[3], 100.0 * best[4] / (n / 16)))
