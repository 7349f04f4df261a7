"""Build libfsb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

No torch headers are involved: the library is plain CUDA C++ behind `extern "C"` entry points
(see include/fsb200.h), so nvcc cross-compiles it in seconds without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libfsb200.so"
STAMP_PATH = PKG_DIR / ".libfsb200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--shared",
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _source_hash() -> str:
    h = hashlib.sha256()
    header = PKG_DIR.parent / "include" / "fsb200.h"
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh"))) + ([header] if header.exists() else []):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc() -> str | None:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if cand and os.path.exists(cand) else None


def is_stale() -> bool:
    if not LIB_PATH.exists() or not STAMP_PATH.exists():
        return True
    return STAMP_PATH.read_text().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every kernel file into one shared library. Raises on any compiler error."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libfsb200.so (set NVCC=/path/to/nvcc)")
    objs = []
    build_dir = PKG_DIR / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in _sources():
        obj = build_dir / (src.stem + ".o")
        cmd = [nvcc, *[f for f in NVCC_FLAGS if f != "--shared"], "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{out}")
        if verbose and out:
            print(out)
        objs.append(str(obj))
    tmp = LIB_PATH.with_suffix(".so.tmp")
    link = [nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp), *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc link failed:\n{r.stdout}")
    os.replace(tmp, LIB_PATH)
    STAMP_PATH.write_text(_source_hash())
    return LIB_PATH


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
