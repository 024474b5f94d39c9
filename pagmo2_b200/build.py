"""Build libpgc.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m pagmo2_b200.build [--force] [--verbose]

Objects go to pagmo2_b200/csrc/_build/, the library to pagmo2_b200/libpgc.so (git-ignored, but it travels to the
GPU box with the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libpgc.so"
OBJDIR = CSRC / "_build"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false: the reference is built for baseline x86-64 (no FMA contraction); the rotation kernels call fma()
# explicitly, everything else keeps the reference's separate multiply/add roundings.
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--threads", "2",
          "-Xptxas", "-v", "-Wno-deprecated-gpu-targets"]


def sources() -> list[Path]:
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cpp")))


def headers() -> list[Path]:
    return sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((HERE.parent / "include").rglob("*.h")))


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in paths:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(COMMON + ARCH).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJDIR.mkdir(parents=True, exist_ok=True)
    hdr_digest = _digest(headers())
    objs, jobs = [], []
    for src in sources():
        obj = OBJDIR / (src.name + ".o")
        stamp = OBJDIR / (src.name + ".stamp")
        want = _digest([src]) + hdr_digest
        objs.append(obj)
        if not force and obj.exists() and stamp.exists() and stamp.read_text() == want:
            continue
        cmd = [NVCC, *ARCH, *COMMON, "-x", "cu", "-c", str(src), "-o", str(obj)]
        jobs.append((src, cmd, stamp, want))

    def run(job):
        src, cmd, stamp, want = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = OBJDIR / (src.name + ".log")
        log.write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        stamp.write_text(want)
        if verbose:
            print(r.stderr, file=sys.stderr)
        return src.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name in ex.map(run, jobs):
                print(f"[pgc build] compiled {name}", file=sys.stderr)
    if jobs or not OUT.exists():
        cmd = [NVCC, *ARCH, "-shared", "-o", str(OUT), *map(str, objs), "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        print(f"[pgc build] linked {OUT}", file=sys.stderr)
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
