"""Builds libmadtp_b200.so in-tree with nvcc for sm_100a (cross-compiles on a machine without a GPU).

    python -m madtp_b200.csrc.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/madtp_b200.h); it links the CUDA runtime statically and
resolves the single driver symbol it needs (cuTensorMapEncodeTiled) at run time, so it loads on CPU-only hosts.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent
REPO = CSRC.parent.parent
LIB = CSRC.parent / "libmadtp_b200.so"
BUILD = CSRC / "build"
SOURCES = ["capi.cu", "gemm.cu", "rowops.cu", "attention.cu", "small_attn.cu", "attn_tc.cu", "cross_attn_tc.cu", "sdft_tc.cu", "dtp.cu", "dtp_apply.cu"]
HEADERS = ["common.cuh", "gemm.cuh", "rowops.cuh", "attention.cuh", "dtp.cuh", "../../include/madtp_b200.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
] + os.environ.get("MADTP_NVCC_EXTRA", "").split()


FASTCALL = CSRC.parent / "_fastcall.so"


def build_fastcall(force: bool = False) -> Path:
    """The CPython trampoline extension (fastcall.c): gcc only, x86-64 only. Optional -- _lib.py falls back to ctypes."""
    import sysconfig
    src = CSRC / "fastcall.c"
    if not force and FASTCALL.exists() and FASTCALL.stat().st_mtime >= src.stat().st_mtime:
        return FASTCALL
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-I", sysconfig.get_paths()["include"], str(src),
           "-o", str(FASTCALL)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"fastcall build failed:\n{r.stdout}\n{r.stderr}")
    return FASTCALL


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        h.update((CSRC / name).read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    stamp = BUILD / "digest.txt"
    digest = _digest()
    build_fastcall(force)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB

    def compile_one(src: str) -> Path:
        obj = BUILD / (src.replace(".cu", ".o"))
        cmd = [NVCC, *FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD / (src + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
    sys.exit(0)
