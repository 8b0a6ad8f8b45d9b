"""In-tree build of libr6dof.so (hand-written CUDA for sm_100a, C ABI of include/r6dof.h).

    python -m rl_rocket_6dof_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.

Staleness is decided by a SHA-256 over every source the library is compiled from (csrc/*.cu, csrc/*.cuh,
include/*.h) plus the compiler flags, stored beside the library (lib/libr6dof.so.srchash) — not by mtimes, which a
checkout, a copy to another box or a touched file make meaningless.  Concurrent builders (one process per GPU under
torchrun all import the package at once) serialise on a file lock, and the library is written to a temporary name
and renamed into place, so nobody can dlopen a half-written file.
"""
from __future__ import annotations

import fcntl
import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libr6dof.so")
HASH_PATH = LIB_PATH + ".srchash"
LOCK_PATH = os.path.join(LIB_DIR, ".build.lock")
SOURCES = [os.path.join(CSRC, "r6_kernels.cu")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def deps() -> list:
    """Everything the library is compiled from."""
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                  glob.glob(os.path.join(os.path.dirname(PKG), "include", "*.h")))


def source_hash() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def find_nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found (needed to build libr6dof.so for sm_100a)")


def built_hash() -> str | None:
    try:
        with open(HASH_PATH) as f:
            return f.read().strip()
    except OSError:
        return None


def needs_build() -> bool:
    return not os.path.exists(LIB_PATH) or built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(LOCK_PATH, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            want = source_hash()
            if not force and os.path.exists(LIB_PATH) and built_hash() == want:
                return LIB_PATH                      # another process built it while we waited for the lock
            tmp = f"{LIB_PATH}.tmp.{os.getpid()}"
            cmd = [find_nvcc()] + NVCC_FLAGS + ["-o", tmp] + SOURCES
            res = subprocess.run(cmd, capture_output=True, text=True)
            log = res.stdout + res.stderr
            with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
                f.write(" ".join(cmd) + "\n" + log)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.unlink(tmp)
                sys.stderr.write(log)
                raise RuntimeError("nvcc failed building libr6dof.so")
            os.replace(tmp, LIB_PATH)
            with open(HASH_PATH + ".tmp", "w") as f:
                f.write(want + "\n")
            os.replace(HASH_PATH + ".tmp", HASH_PATH)
            if verbose:
                print(log)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
