"""Builds ``libgsb.so`` (the C-ABI CUDA library, sm_100a only) in-tree with nvcc.

    python -m gaussiansplattingmlx_b200.build [--force]

nvcc cross-compiles without a GPU.  ``project.cu`` (forward only) and ``adam.cu`` are compiled with
``--fmad=false`` so that their arithmetic rounds exactly like the ``-ffp-contract=off`` CPU oracle
(bit-exact projection geometry → bit-exact tile lists; bit-exact Adam); both are HBM-bound, the
lost FMA contraction is not on their critical path.  Everything else is compiled with FMA.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
BUILD = HERE / "csrc" / "build"
LIB = HERE / "libgsb.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"]
SOURCES = {
    "project.cu": ["--fmad=false"],
    "project_bwd.cu": [],
    "adam.cu": ["--fmad=false"],
    "densify.cu": ["--fmad=false"],
    "binning.cu": [],
    "tilelists.cu": [],
    "raster.cu": [],
    "loss.cu": [],
    "api.cu": [],
}


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [HERE.parent / "include" / "gsb.h", Path(__file__)]
    return max(p.stat().st_mtime for p in hdrs)


def _compile(name: str, extra, force: bool, log_dir: Path) -> Path:
    src = CSRC / name
    obj = BUILD / (name + ".o")
    newest = max(src.stat().st_mtime, _deps_mtime())
    if not force and obj.exists() and obj.stat().st_mtime >= newest:
        return obj
    cmd = [NVCC, *ARCH, *COMMON, *extra, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (log_dir / (name + ".ptxas.log")).write_text(r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed for {name}")
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(parents=True, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda kv: _compile(kv[0], kv[1], force, BUILD), SOURCES.items()))
    if force or not LIB.exists() or any(o.stat().st_mtime > LIB.stat().st_mtime for o in objs):
        cmd = [NVCC, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
