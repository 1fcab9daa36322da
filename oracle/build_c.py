"""Builds oracle/c/libmamri_oracle.so (plain-C oracle; TEST INFRASTRUCTURE, see oracle/__init__.py).
Tries OpenMP first and falls back to a single-threaded build if no OpenMP-capable gcc is found."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent / "c"
SRC = HERE / "mamri_oracle.c"
LIB = HERE / "libmamri_oracle.so"
BASE = ["-O3", "-march=x86-64-v3", "-fPIC", "-fvisibility=hidden", "-std=gnu11", "-Wall", "-Wextra", "-Wno-unknown-pragmas"]


def build(force: bool = False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    errors = []
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            cmd = [cc, *BASE, *omp, "-shared", "-o", str(LIB), str(SRC), "-lm"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode == 0:
                return LIB
            errors.append(f"{' '.join(cmd)}\n{r.stderr}")
    raise RuntimeError("could not build the C oracle:\n" + "\n".join(errors))


if __name__ == "__main__":
    print(build(force=True))
