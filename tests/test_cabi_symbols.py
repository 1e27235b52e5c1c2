"""CPU-side checks of the product library: it loads, exports every symbol include/pms.h declares,
and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

from helpers import cuda_available
from pyminisim_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pms.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pms_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pyminisim_b200 import build
    build.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pms.h but not exported"
    # the Python binding covers the same surface
    assert set(declared) == set(_capi.PROTOTYPES)


def test_defaults_match_reference_constants():
    lib = _capi.load()
    p = _capi.HsfmParamsC()
    lib.pms_default_hsfm_params(ctypes.byref(p))
    # HSFMParams.create_default (_hsfm.py:31-42), ctor defaults (:198-208), core/_constants.py
    assert (p.tau, p.A, p.B, p.k_1, p.k_2, p.k_o, p.k_d, p.k_lambda, p.alpha) == (0.2, 2000.0, 0.08, 1.2e5, 2.4e5, 1.0, 500.0, 0.3, 3.0)
    assert (p.pedestrian_mass, p.pedestrian_radius, p.robot_radius) == (70.0, 0.3, 0.35)
    o = _capi.OrcaParamsC()
    lib.pms_default_orca_params(ctypes.byref(o))
    assert (o.neighbor_dist, o.max_neighbors, o.time_horizon, o.radius, o.goal_reaching_threshold) == (1.0, 10, 2.0, 0.3, 0.2)


@pytest.mark.skipif(cuda_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    from pyminisim_b200.batched import BatchedSimulation
    with pytest.raises(_capi.PmsError) as ei:
        BatchedSimulation(2, 4)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pyminisim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "pms_oracle", "liboracle", "orc_"):
                    assert needle not in src, f"{f} references the oracle ({needle})"


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """include/pms.h is the boundary a maintainer binds from C / cgo / JNI: it must compile as C99 without CUDA or C++
    headers, and a C caller must link against libpms_b200.so.  The program only queries defaults and the error path of
    pms_create (no device here), i.e. no compute without a GPU."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    from pyminisim_b200 import build
    build.build()
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stdio.h>
#include "pms.h"
int main(void) {
    pms_hsfm_params p;
    pms_default_hsfm_params(&p);
    if (p.A != 2000.0 || p.B != 0.08) return 2;
    pms_sim *h = 0;
    int rc = pms_create(0, 4, 8, PMS_FP32, &p, &h);
    if (rc == PMS_OK) { printf("created\n"); return pms_destroy(h); }   /* GPU box */
    if (rc != PMS_ERR_CUDA || h != 0) return 3;                          /* CPU box: loud failure, no fallback */
    printf("%s\n", pms_last_error());
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_capi.LIB_PATH)
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libpms_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "created" in out.stdout or "CUDA" in out.stdout or "fallback" in out.stdout


def test_philox_known_answers():
    """The device noise generator is Philox4x32-10: Random123's published known-answer vectors (kat_vectors: philox4x32 10)."""
    lib = _capi.load()
    u32x4, u32x2 = ctypes.c_uint32 * 4, ctypes.c_uint32 * 2
    pi_ctr = (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)
    pi_key = (0xa4093822, 0x299f31d0)
    for ctr, key, want in (((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
                           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
                           (pi_ctr, pi_key, (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))):
        out = u32x4()
        lib.pms_philox4x32_10(u32x4(*ctr), u32x2(*key), out)
        assert tuple(out) == want, [hex(v) for v in out]
