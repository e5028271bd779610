#!/usr/bin/env python
"""Per-CUDA-source-line summary of an ncu report (needs -lineinfo + --import-source on):
   python tools/ncu_lines.py report.ncu-rep 'kernel substring' [top]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fn = fp = None
hdr = None
acc = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fp = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        fn = r[1]; hdr = None; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or pat not in (fn or "") or len(r) < len(hdr):
        continue
    if r[2] != "-":      # SASS row under a source line; the line row (Address '-') already aggregates
        continue
    d = dict(zip(hdr[4:], r[4:]))
    key = (fp, int(r[0]))
    a = acc.setdefault(key, [r[1], 0, 0])
    a[1] += int(d.get("# Samples", 0) or 0)
    a[2] += int(d.get("Instructions Executed", 0) or 0)
ts = sum(a[1] for a in acc.values()); ti = sum(a[2] for a in acc.values())
print("kernel~%s: %d samples, %d warp instructions" % (pat, ts, ti))
for (f, ln), a in sorted(acc.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.2f%% smp %5.2f%% inst  %s:%d  %s" % (100.0 * a[1] / max(ts, 1), 100.0 * a[2] / max(ti, 1), f, ln, a[0].strip()[:110]))
