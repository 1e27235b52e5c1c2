"""Registers / spills per kernel from an `nvcc -Xptxas=-v` log:  python tools/ptxas_table.py build.log [other.log]
With two logs, prints only the entries that differ."""
import re
import subprocess
import sys


def parse(path):
    out, name = {}, None
    for line in open(path, errors="replace"):
        m = re.search(r"Compiling entry function '([^']+)'", line) or re.search(r"Function properties for (\S+)", line)
        if m:
            name = m.group(1)
            out.setdefault(name, {})
            continue
        if name is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            out[name].update(stack=int(m.group(1)), st=int(m.group(2)), ld=int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            out[name]["regs"] = int(m.group(1))
    return out


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:110]
    except Exception:
        return n


a = parse(sys.argv[1])
b = parse(sys.argv[2]) if len(sys.argv) > 2 else None
for k in sorted(a):
    if b is None:
        print(demangle(k), a[k])
    elif a.get(k) != b.get(k):
        print(demangle(k), a.get(k), "->", b.get(k))
if b:
    for k in sorted(set(b) - set(a)):
        print("new:", demangle(k), b[k])
