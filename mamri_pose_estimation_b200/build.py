"""Builds ``libmamri_b200.so`` in-tree with nvcc for sm_100a (B200 only; no other
architecture, no JIT).  ``python -m mamri_pose_estimation_b200.build`` or
``__graft_entry__.build()``."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libmamri_b200.so"
TRACE_LIB = HERE / "libmamri_b200_trace.so"
SOURCES = ["segment.cu", "ccl.cu", "stats.cu", "entry.cu", "surface.cu", "pose.cu", "phantom.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> Path:
    """Compiles the translation units (in parallel) and links the shared library.  ``trace=True`` builds the
    diagnostic variant ``libmamri_b200_trace.so`` (-DMAMRI_KTRACE: every kernel stamps %globaltimer, see
    common.cuh and tools/ktrace.py); the product never loads it unless MAMRI_LIB points at it."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    headers = [CSRC / "common.cuh", HERE.parent / "include" / "mamri_b200.h"]
    lib = TRACE_LIB if trace else LIB
    suffix = ".trace.o" if trace else ".o"
    extra = ["-DMAMRI_KTRACE"] if trace else []
    objs, jobs = [], []
    for src in SOURCES:
        s = CSRC / src
        o = CSRC / (s.stem + suffix)
        objs.append(o)
        if force or _stale(o, [s, *headers]):
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(s), "-o", str(o)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for r in ex.map(lambda c: subprocess.run(c, check=False, capture_output=True, text=True), jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed: " + " ".join(r.args))
    if force or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(lib), *map(str, objs)]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv))
