"""Append selected metrics of an `ncu --set full` report to profiles/r01_ncu_render_kernels.csv.
Usage: python profiles/ncu_summary.py <report.ncu-rep> <version-tag>"""
import csv
import io
import os
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r01_ncu_render_kernels.csv")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
cols = next(csv.reader(open(out)))
with open(out, "a", newline="") as f:
    w = csv.writer(f)
    for r in rows[2:]:
        w.writerow([tag] + [r[hdr.index(c)] if c in hdr else "" for c in cols[1:]])
print("appended", len(rows) - 2, "rows to", out)
