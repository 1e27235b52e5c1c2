"""Turn an `ncu --set full --import-source on` report into the markdown summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "<title>" "<capture command>" [units-per-launch] > profiles/xx.md
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --json kernel=hsfm_warp_kernel worlds=4096 peds=64 sim_steps=1000 \
        precision=fp32 source=profiles/xx.md      # appends the record bench.py reads to profiles/ncu_captures.json

Reads the raw page (pipe utilisation, stalls, DRAM traffic, occupancy limits) and the source page (executed
warp-instructions per opcode; `units-per-launch`, e.g. worlds x steps, turns them into a per-unit count).
Runs here on the CPU box (ncu -i needs no GPU)."""
import csv
import subprocess
import sys
from collections import Counter

RAW = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (shared-memory limit)"),
    ("launch__cluster_size", "cluster size"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe: instructions % of peak"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe: cycles active % (packed FFMA2 = 2 cycles)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % (shared memory + shuffles)"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]
STALLS = ["wait", "barrier", "short_scoreboard", "math_pipe_throttle", "not_selected", "long_scoreboard", "mio_throttle",
          "no_instruction", "dispatch_stall", "branch_resolving", "lg_throttle", "membar"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def to_bytes(unit, val):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(val) * scale[unit]


def json_record():
    import json
    import os
    rep = sys.argv[1]
    kv = dict(a.split("=", 1) for a in sys.argv[3:])
    rows = page(rep, "raw")
    m = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    rec = {"kernel": kv["kernel"], "worlds": int(kv["worlds"]), "peds": int(kv["peds"]), "sim_steps": int(kv["sim_steps"]),
           "precision": kv["precision"], "source": kv.get("source", rep), "kernel_name": m["Kernel Name"][1],
           "duration_ms": float(m["gpu__time_duration.sum"][1]) * {"ms": 1.0, "us": 1e-3, "s": 1e3}[m["gpu__time_duration.sum"][0]],
           "dram_bytes": int(to_bytes(*m["dram__bytes_read.sum"]) + to_bytes(*m["dram__bytes_write.sum"])),
           "fma_pipe_cycles_pct": float(m["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][1]),
           "xu_pipe_pct": float(m["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"][1]),
           "issue_slots_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][1])}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_captures.json")
    recs = json.load(open(path)) if os.path.exists(path) else []
    recs.append(rec)
    json.dump(recs, open(path, "w"), indent=1)
    print(json.dumps(rec))


def main():
    if len(sys.argv) > 2 and sys.argv[2] == "--json":
        return json_record()
    rep, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    units = float(sys.argv[4]) if len(sys.argv) > 4 else None
    rows = page(rep, "raw")
    hdr, unit, val = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, unit, val)}
    print(f"# {title}\n")
    print(f"Capture: `{cmd}` on one B200 (sm_100a), driver 580.159, CUDA 12.9; the report itself stays out of git "
          f"(`{rep}`).\n")
    kn = m.get("Kernel Name", ("", ""))[1]
    if kn:
        print(f"Kernel: `{kn}`\n")
    print("| metric | value | unit |\n|---|---|---|")
    for key, label in RAW:
        if key in m:
            print(f"| {label} (`{key}`) | {m[key][1]} | {m[key][0]} |")
    for s in STALLS:
        key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if key in m and float(m[key][1]) >= 0.05:
            print(f"| stall: {s.replace('_', ' ')} (`{key}`) | {float(m[key][1]):.3f} | warps / issue |")
    src = page(rep, "source")
    if len(src) > 2:
        h = src[1]
        ix = {n: i for i, n in enumerate(h)}
        ops, total = Counter(), 0
        for r in src[2:]:
            try:
                n = int(r[ix["Instructions Executed"]])
            except (ValueError, IndexError):
                continue
            s = r[ix["Source"]].split()
            if not s:
                continue
            op = s[1] if s[0].startswith("@") and len(s) > 1 else s[0]
            ops[op.split(".")[0]] += n
            total += n
        print(f"\nExecuted warp-instructions by opcode (source page; total {total:,}"
              + (f", {total / units:.1f} per unit of {units:,.0f} units in this launch" if units else "") + "):\n")
        print("| opcode | executed |" + (" per unit |" if units else "") + "\n|---|---|" + ("---|" if units else ""))
        for op, n in ops.most_common(22):
            print(f"| {op} | {n:,} |" + (f" {n / units:.1f} |" if units else ""))


if __name__ == "__main__":
    main()
