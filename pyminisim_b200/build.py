"""Build libpms_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpms_b200.so")
SOURCES = ["pms_api.cu", "hsfm_warp.cu", "hsfm_duo.cu", "replay.cu", "orca.cu"]   # orca.cu is compiled with -fmad=false (RVO2 rounds every op)
HEADERS = ["pms_common.cuh", "hsfm_small.cuh", "hsfm_warp.cuh", "hsfm_warp_api.cuh", "hsfm_duo.cuh", "hsfm_duo_api.cuh", "hsfm_large.cuh", "sensors.cuh", "sensor_fires.cuh", "replay_api.cuh", "orca.cuh"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libpms_b200.so")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(os.path.dirname(HERE), "include", "pms.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    if not force and out == LIB and not is_stale():
        return LIB
    nvcc = nvcc_path()
    common = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"] + \
             [f"-D{d}" for d in defines]
    if verbose:
        common.insert(0, "-Xptxas=-v")
    objs, procs = [], []
    for src in SOURCES:   # one nvcc per translation unit, in parallel
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + (".o" if out == LIB else ".exp.o"))
        extra = ["-fmad=false"] if src == "orca.cu" else []
        procs.append((src, subprocess.Popen([nvcc] + common + extra + ["-c", os.path.join(CSRC, src), "-o", obj])))
        objs.append(obj)
    for src, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, f"nvcc {src}")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs, check=True)
    return out


if __name__ == "__main__":
    import sys
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(defs), verbose="-v" in sys.argv, defines=defs,
                out=outs[0] if outs else LIB))
