"""Markdown table from the JSON lines of tools/parity_report.py:  python tools/parity_md.py parity.jsonl >> profiles/rN_parity.md"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
print("| scenario | N | precision | steps | worlds | gpu-only blow-ups | position error max | position error median | velocity error max |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r['scenario']} | {r['n']} | {r['precision']} | {r['steps']} | {r['worlds']} | {r['blown_up_on_gpu_only']} | "
          f"{r['pos_max']:.2e} | {r['pos_median']:.2e} | {r['vel_max']:.2e} |")
