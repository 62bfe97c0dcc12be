"""The C-ABI shared library: builds for sm_100a, loads, exports every declared symbol, and fails
loudly (no CPU fallback) when no CUDA device is present.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from zodipy_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _cabi.load()


def test_header_symbols_all_exported(lib):
    header = open(os.path.join(ROOT, "include", "zodi_b200.h")).read()
    declared = set(re.findall(r"\b(zodi_[a-z0-9_]+)\s*\(", header))
    declared -= {"zodi_model_s"}
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)


def test_struct_layouts_match_header_sizes(lib):
    # sizes computed from the header's field lists (8-byte alignment throughout)
    assert C.sizeof(_cabi.ComponentDesc) == 8 + 3 * 8 + 4 * 8 + 8 * 8 + 2 * 8 + 2 * 8 + 2 * 8
    assert C.sizeof(_cabi.ModelDesc) == 6 * 4 + 7 * 8 + 4 * 8 + 16 * C.sizeof(_cabi.ComponentDesc)
    assert C.sizeof(_cabi.EvalArgs) == 8 + 16 + 24 + 24 + 8 + 16 + 16 + 8 + 8 + 8 * 8 + 16 + 16 + 16
    assert lib.zodi_abi_version() == _cabi.ABI_VERSION


def test_struct_layouts_match_the_c_compiler(tmp_path):
    """sizeof / offsetof of every ABI struct as gcc lays include/zodi_b200.h out == the ctypes mirror."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc unavailable")
    structs = {"zodi_component_desc": _cabi.ComponentDesc, "zodi_model_desc": _cabi.ModelDesc,
               "zodi_eval_args": _cabi.EvalArgs, "zodi_ephemeris_desc": _cabi.EphemerisDesc,
               "zodi_healpix_args": _cabi.HealpixArgs, "zodi_lonlat_args": _cabi.LonLatArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "zodi_b200.h"', "int main(void) {"]
    for cname, ctype in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for field, _ in ctype._fields_:
            lines.append(f'  printf("{cname} {field} %zu\\n", offsetof({cname}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, field, value = line.split()
        ctype = structs[cname]
        expect = C.sizeof(ctype) if field == "size" else getattr(ctype, field).offset
        assert int(value) == expect, (cname, field, value, expect)
        seen += 1
    assert seen == sum(len(t._fields_) + 1 for t in structs.values())
    assert _cabi.MAX_COMPS == 16 and _cabi.MAX_PEERS == 8


def test_sass_is_sm100a_only():
    import subprocess

    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_silent_cpu_fallback(lib):
    """Without a GPU every compute entry point must fail with a CUDA error, never compute."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    n = C.c_int(-1)
    assert lib.zodi_device_count(C.byref(n)) == -2  # ZODI_ERR_CUDA
    assert b"cuda" in lib.zodi_last_error().lower()
    import zodipy_b200 as zp

    m = zp.Model(zp.Quantity(25, "um"))
    import numpy as np

    with pytest.raises(_cabi.ZodiError):
        m.evaluate_xyz(np.array([[1.0], [0.0], [0.0]]), np.array([0.0, 1.0, 0.0]))
    out = C.c_double(0)
    assert lib.zodi_peak_probe(0, 0, C.byref(out)) == -2


def test_invalid_descriptors_are_rejected(lib):
    import numpy as np

    import zodipy_b200 as zp
    from zodipy_b200.spec import pack_desc

    spec = zp.Model(zp.Quantity(25, "um")).spec
    handle = C.c_void_p()
    desc, keep = pack_desc(spec)
    desc.abi_version = 99
    assert lib.zodi_model_create(C.byref(desc), 0, C.byref(handle)) == -1
    desc, keep = pack_desc(spec)
    desc.n_nodes = 0
    assert lib.zodi_model_create(C.byref(desc), 0, C.byref(handle)) == -1
    bad = dict(spec)
    bad["table"] = np.array([np.geomspace(40, 550, 100), spec["table"][1]])
    desc, keep = pack_desc(bad)
    assert lib.zodi_model_create(C.byref(desc), 0, C.byref(handle)) == -4  # non-uniform knots
    assert b"uniformly" in lib.zodi_last_error()
    assert lib.zodi_model_create(None, 0, C.byref(handle)) == -1
    assert lib.zodi_evaluate(None, None) == -1
