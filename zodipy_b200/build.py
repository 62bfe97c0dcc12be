"""Build the CUDA library in-tree: ``python -m zodipy_b200.build`` -> zodipy_b200/libzodi_b200.so.

sm_100a only (B200).  nvcc cross-compiles without a GPU.  The kernel families are separate translation
units (csrc/zodi_launch_*.cu, some compiled once per arithmetic type / lane count) built in parallel
and linked into ONE shared library.  The .so is git-ignored but travels to the GPU box with the gpurun
snapshot; objects are cached under zodipy_b200/build/ (git-ignored).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libzodi_b200.so")
HEADERS = ["zodi_device.cuh", "zodi_fp64_tables.cuh", "zodi_kernels.cuh", "zodi_kelsall.cuh", "zodi_kelsall_x2.cuh",
           "zodi_multiband.cuh", "zodi_rrm.cuh", "zodi_rrm_x2.cuh", "zodi_misc_kernels.cuh", "zodi_model_build.hpp", "zodi_launch.hpp",
           os.path.join("..", "..", "include", "zodi_b200.h")]

_TYPES = (("f32", "float"), ("f64", "double"))
# (object name, source, extra defines)
UNITS = [("capi", "zodi_capi.cu", [])]
for _suffix, _real in _TYPES:
    for _fam in ("generic", "kelsall", "multiband", "rrm"):
        UNITS.append((f"{_fam}_{_suffix}", f"zodi_launch_{_fam}.cu",
                      [f"-DZODI_TU_REAL={_real}", f"-DZODI_TU_SUFFIX={_suffix}", f"-DZODI_TU_IS_{_suffix.upper()}"]))
for _lanes in (1, 2, 4, 8):
    UNITS.append((f"x2_l{_lanes}", "zodi_launch_x2.cu", [f"-DZODI_TU_LANES={_lanes}"]))

UNITS = [u for u in UNITS if os.path.exists(os.path.join(CSRC, u[1]))]
HEADERS = [h for h in HEADERS if os.path.exists(os.path.join(CSRC, h))]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC"]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the CUDA library cannot be built")
    return nvcc


def _deps():
    srcs = sorted({u[1] for u in UNITS})
    return [os.path.join(CSRC, s) for s in srcs + HEADERS] + [os.path.abspath(__file__)]


def up_to_date(lib_path: str = LIB_PATH) -> bool:
    if not os.path.exists(lib_path):
        return False
    t = os.path.getmtime(lib_path)
    return all(os.path.getmtime(d) <= t for d in _deps())


def _headers_digest(extra) -> str:
    h = hashlib.sha1()
    for name in HEADERS:
        with open(os.path.join(CSRC, name), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + list(extra)).encode())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False, defines=(), lib_path: str = LIB_PATH) -> str:
    """Compile (objects whose source, headers and flags are unchanged are reused) and link.

    ``defines``: extra ``-D...`` flags for A/B builds of kernel variants into another ``lib_path``
    (loaded through the ``ZODI_B200_LIB`` environment variable, see ``_cabi``).
    """
    defines = list(defines)
    if not force and not defines and up_to_date(lib_path):
        return lib_path
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    tag = _headers_digest(defines)

    def compile_unit(unit):
        name, src, extra = unit
        src_path = os.path.join(CSRC, src)
        with open(src_path, "rb") as fh:
            src_tag = hashlib.sha1(fh.read() + tag.encode() + " ".join(extra).encode()).hexdigest()[:16]
        obj = os.path.join(OBJ_DIR, f"{name}.{src_tag}.o")
        if os.path.exists(obj) and not force:
            return obj, ""
        for old in os.listdir(OBJ_DIR):  # one cached object per unit and variant tag is enough
            if old.startswith(name + ".") and old.endswith(".o") and not defines:
                os.remove(os.path.join(OBJ_DIR, old))
        cmd = [nvcc, *NVCC_FLAGS, *defines, *extra, "-c", src_path, "-o", obj + f".tmp{os.getpid()}"]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src} {extra}:\n{proc.stdout}{proc.stderr}")
        os.replace(obj + f".tmp{os.getpid()}", obj)
        return obj, proc.stdout + proc.stderr

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_unit, UNITS))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    # link next to the target and rename: a reader (another process, a gpurun snapshot) never
    # sees a half-written library
    tmp = lib_path + f".tmp{os.getpid()}"
    proc = subprocess.run([nvcc, *LINK_FLAGS, "-o", tmp, *[obj for obj, _ in results]], capture_output=True, text=True)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    os.replace(tmp, lib_path)
    return lib_path


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    out = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs,
                lib_path=os.path.abspath(out[0]) if out else LIB_PATH))
