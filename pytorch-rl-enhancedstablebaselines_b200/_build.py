"""In-tree nvcc build of the C-ABI library (``libcstr_b200.so``) for sm_100a.

The library is plain CUDA C++ behind ``extern "C"`` (include/cstr_b200.h) — no torch headers — so a
single ``nvcc -shared`` produces it in seconds and the ``.so`` travels with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from typing import List

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libcstr_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-shared",
    # link the CUDA runtime dynamically (the same libcudart.so.12 torch has loaded, else the toolkit's): the artefact then carries only
    # the runtime entry points this code calls instead of a private copy of the whole runtime and its symbol table
    "-cudart", "shared",
    "-Xlinker", "-rpath=/usr/local/cuda/lib64",
]


def sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps() -> List[str]:
    inc = os.path.join(os.path.dirname(PKG_DIR), "include", "cstr_b200.h")
    return sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [inc]


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in _deps())


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built on this host")


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu into libcstr_b200.so (cross-compiles without a GPU).  Safe when several processes of one job call it
    at once (``torchrun`` ranks that all find the library stale): one builds under a file lock, the others wait and find it fresh."""
    if not force and not is_stale():
        return LIB_PATH
    import fcntl

    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():  # another process built it while this one waited
                return LIB_PATH
            extra = os.environ.get("CSTR_NVCC_EXTRA", "").split()  # e.g. -DCSTR_TC_TIMING / -DCSTR_TD3_TIMING (clock64 phase instrumentation)
            tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
            cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-o", tmp, *sources()]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            # the image's $CC wrapper is not a usable nvcc host compiler; let nvcc pick the system g++
            env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
            res = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
