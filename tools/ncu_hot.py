#!/usr/bin/env python
"""Hot spots of a kernel from an ncu report: `ncu_hot.py report.ncu-rep [top N]` -- per-SASS-instruction stall samples
(source page, SASS view) ranked, with the dominant stall reason and the executed count."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    print(f"instructions {len(rows)}, samples {tot}")
    agg = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
    print("stall totals:", ", ".join(f"{c[6:]} {v}" for c, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    ranked = sorted(enumerate(rows), key=lambda ir: -int(ir[1]["# Samples"] or 0))[:top]
    for i, r in sorted(ranked):
        n = int(r["# Samples"] or 0)
        why = max(stall_cols, key=lambda c: int(r[c] or 0))
        print(f"{i:5d} {n:6d} {100.0 * n / tot:5.1f}%  x{r['Instructions Executed']:>8s}  {why[6:]:12s} {r['Source'].strip()}")


if __name__ == "__main__":
    main()
