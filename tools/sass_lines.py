#!/usr/bin/env python
"""Line-annotated SASS of one kernel of a cubin / object: `sass_lines.py file.o kernel-substring [src-file-substring]`.
Prints `srcline | address opcode operands` (the -lineinfo mapping of nvdisasm -g), for reading hot loops without a GPU."""
import re
import subprocess
import sys
import tempfile
import os


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    src_pat = sys.argv[3] if len(sys.argv) > 3 else ""
    tmp = tempfile.mkdtemp()
    if not obj.endswith(".cubin"):
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    else:
        cubins = [obj]
    for cb in cubins:
        out = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
        on, cur = False, ""
        for ln in out.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                on = pat in m.group(1)
                if on:
                    print("==== " + m.group(1))
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                f = os.path.basename(m.group(1))
                cur = f"{f}:{m.group(2)}"
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
            if m and (not src_pat or src_pat in cur):
                print(f"{cur:24s} | {m.group(1)} {m.group(2)}")
            elif re.match(r"\s*\.L_", ln):
                print(ln.strip())


if __name__ == "__main__":
    main()
