"""Build libreach_sm100a.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libreach_sm100a.so"
EXTRA = os.environ.get("REACH_NVCC_EXTRA", "").split()     # e.g. -DREACH_CHAIN_PROFILE (diagnostics builds)
NVCC_FLAGS = EXTRA + [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (cand == "nvcc" or Path(cand).exists()):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(CSRC.glob("*.cu"))


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu to an object and link the shared library. Incremental by mtime."""
    nvcc = _nvcc()
    objdir = PKG_DIR / "build"
    objdir.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list((PKG_DIR.parent / "include").glob("*.h"))
    hdr_mtime = max(h.stat().st_mtime for h in headers)
    objs, relink = [], force or not LIB_PATH.exists()
    for src in sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_mtime):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
            relink = True
    if relink or any(o.stat().st_mtime > LIB_PATH.stat().st_mtime for o in objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *map(str, objs),
               "-Xlinker", "--exclude-libs=ALL"]
        subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
