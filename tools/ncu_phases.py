"""Stall samples of a kernel by windows of SASS instructions (address order): `ncu_phases.py report.ncu-rep [window] [units]`.
units = launches of the loop body (e.g. worlds x steps): shows executed instructions per unit in each window."""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]; win = int(sys.argv[2]) if len(sys.argv) > 2 else 100; units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
print(f"instructions {len(rows)}, samples {tot}")
for w0 in range(0, len(rows), win):
    blk = rows[w0:w0 + win]
    n = sum(int(r["# Samples"] or 0) for r in blk)
    if n < tot * 0.002:
        continue
    ex = sum(int(r["Instructions Executed"] or 0) for r in blk) / units
    ops = Counter((r["Source"].split()[1] if r["Source"].split()[0].startswith("@") else r["Source"].split()[0]).split(".")[0] for r in blk if r["Source"].split())
    st = Counter()
    for r in blk:
        for c in stall_cols:
            st[c[6:]] += int(r[c] or 0)
    top = ", ".join(f"{k} {100 * v / n:.0f}%" for k, v in st.most_common(4))
    print(f"{w0:5d}-{w0 + len(blk) - 1:5d} {100.0 * n / tot:5.1f}% samples  {ex:7.1f} inst/unit  {n / max(ex * units, 1) * 1e3:6.1f} samples/kinst | {top} | " +
          " ".join(f"{k}:{v}" for k, v in ops.most_common(6)))
