"""Builds libznippy_cuda.so (sm_100a) in-tree with nvcc.  `python -m znippy_b200.build [--force] [-v]`.

The library is five translation units (C ABI + decode/hash kernels, compression kernels, the device-wide zstd decode
pipeline, the native container, the envelope layer); each is compiled to an object only when one of the files it includes changed, the
objects are compiled in parallel and linked into one shared library."""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
SO = os.path.join(HERE, "libznippy_cuda.so")
UNITS = ["znippy_cuda.cu", "compress_tu.cu", "zpipe_tu.cu", "container.cpp", "envelope.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libznippy_cuda.so cannot be built (there is no CPU fallback)")


def _deps(path: str, seen: set[str] | None = None) -> set[str]:
    """`path` and every file it includes with quotes, recursively."""
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path, encoding="utf-8", errors="replace") as f:
        for inc in _INC.findall(f.read()):
            _deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def sources() -> list[str]:
    out: set[str] = set()
    for u in UNITS:
        out |= _deps(os.path.join(CSRC, u))
    return sorted(out)


def _obj(unit: str) -> str:
    return os.path.join(OBJ, os.path.splitext(unit)[0] + ".o")


def _unit_stale(unit: str) -> bool:
    o = _obj(unit)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(os.path.getmtime(s) > t for s in _deps(os.path.join(CSRC, unit)))


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(_unit_stale(u) or os.path.getmtime(_obj(u)) > t for u in UNITS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or stale()):
        return SO
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_unit(unit: str):
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, unit), "-o", _obj(unit)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        if os.environ.get("ZN_TRACE_BUILD"):  # development: phase counters in the pipeline's exec kernel
            cmd.insert(1, "-DZP_TRACE")
        return unit, subprocess.run(cmd, capture_output=True, text=True)

    todo = [u for u in UNITS if force or _unit_stale(u)]
    with ThreadPoolExecutor(max_workers=max(1, len(todo))) as ex:
        for unit, r in ex.map(compile_unit, todo):
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {unit}:\n" + r.stdout + r.stderr)
            if verbose:
                print(r.stderr)
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, *[_obj(u) for u in UNITS]],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
