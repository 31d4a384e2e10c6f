#!/bin/bash
# tools/ncu_export.sh REPORT.ncu-rep OUTPREFIX [KERNEL-REGEX] -- on the GPU box: turn a (large) report into the two small CSV pages
# we read back home (gpurun merges at most 64 MiB): raw metrics (every kernel in the report) and the per-source-line page
# (of the kernels that match KERNEL-REGEX, default: all)
ncu -i "$1" --page raw --csv > "$2_raw.csv" 2>/dev/null
FILTER=()
[ -n "$3" ] && FILTER=(-k "regex:$3")
ncu -i "$1" "${FILTER[@]}" --page source --print-source sass,cuda --csv 2>/dev/null | python3 -c "
import sys, csv
# keep the per-CUDA-line rows only (rows whose first column is a line number) plus the header rows
w = csv.writer(sys.stdout)
for r in csv.reader(sys.stdin):
    if r and (r[0] in ('File Path', 'Function Name', 'Line No') or r[0].isdigit()):
        w.writerow(r[:12])
" > "$2_lines.csv"
