"""Builds libdqlb200.so in-tree with nvcc for sm_100a (no JIT, no torch extension machinery)."""
from __future__ import annotations

import os
import pathlib
import shutil
import subprocess
import sys

PKG = pathlib.Path(__file__).resolve().parent
SRC = PKG / "csrc" / "dqlb200.cu"
LIB = PKG / "libdqlb200.so"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--fmad=false",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    newest = max(p.stat().st_mtime for p in list((PKG / "csrc").glob("*")) + [PKG.parent / "include" / "dqlb200.h"])
    return newest > LIB.stat().st_mtime


def build(force: bool = False, verbose: bool = False) -> pathlib.Path:
    if not force and not needs_build():
        return LIB
    out = pathlib.Path(os.environ.get("DQL_BUILD_OUT") or LIB)          # DQL_BUILD_OUT: kernel A/B variants (build_variants/)
    tmp = out.with_suffix(".so.tmp")
    cmd = [nvcc_path(), *NVCC_FLAGS, *os.environ.get("DQL_NVCC_EXTRA", "").split(), "-o", str(tmp), str(SRC)]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, out)                # atomic: a concurrent reader (gpurun snapshot) never sees a half-written library
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
