#!/usr/bin/env python3
"""tools/one_scan.py GiB PATTERN [k=2 nocase=1 ... list=1 reps=3] -- one pattern over a device-resident synthetic corpus, a few
times: the thing to put under `ncu -k regex:... -c 1`.  Development tool, not part of the driver contract."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import agrep_b200 as ag

gib, pat = float(sys.argv[1]), sys.argv[2]
kw = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    kw[k] = v if k == "delim" else int(v)
reps, want_list, para, ordinals = kw.pop("reps", 3), kw.pop("list", 0), kw.pop("para", 0), kw.pop("ordinals", 0)
n = int(gib * (1 << 30)) // 4096 * 4096
t = torch.empty(n + 4096, dtype=torch.uint8, device="cuda")
t[n:].zero_()
ag.corpus_device(t.data_ptr(), n, needle=pat if len(pat) < 60 else "", needle_every=4096, needle_maxedits=3, paragraphs=bool(para))
torch.cuda.synchronize()
p = ag.Pattern(pat, **kw)
cap = 1 << 22
rec = torch.empty((cap, 4), dtype=torch.int64, device="cuda") if want_list else None
for _ in range(reps):
    r = p.scan_device(t.data_ptr(), n, d_records=rec.data_ptr() if want_list else 0, capacity=cap if want_list else 0, ordinals=bool(ordinals))
    print("front %.3f ms  rest %.3f ms  %.1f GB/s  matched %d flagged %d" % (r.ms_front, r.ms_records, n / (r.ms_front + r.ms_records) / 1e6, r.n_matched, r.n_flagged))
