"""Builds the C-ABI library ``coral_b200/lib/libcoral_b200.so`` in-tree with nvcc for sm_100a.

No torch types cross the boundary, so this is a plain ``nvcc -shared`` build (static
cudart); the .so travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libcoral_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

CU_SOURCES = ["lm.cu", "beam.cu", "beam_frames.cu", "greedy.cu", "edit.cu", "text.cu"]
CC_SOURCES = ["lm_host.cc", "normalise.cc"]
HEADERS = ["lm_tables.h", "lm_host.h", "beam_core.h", "beam_launch.cuh", "handles.h", "common.cuh", os.path.join("..", "..", "include", "coral_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--fmad=false", "-I", OBJ_DIR,
]


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the coral_b200 CUDA library cannot be built")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _unicode_tables() -> str:
    """CPython's Unicode behaviour as C tables for the text normaliser (csrc/normalise.cc):
    generated from the running interpreter, so the two cannot drift apart."""
    gen = os.path.join(CSRC, "gen_unicode_tables.py")
    out = os.path.join(OBJ_DIR, "unicode_tables.h")
    if _stale(out, [gen]):
        import runpy

        runpy.run_path(gen)["main"](out + ".tmp")
        os.replace(out + ".tmp", out)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    _unicode_tables()
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    objs = []
    for src in CU_SOURCES + CC_SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ_DIR, src + ".o")
        objs.append(op)
        if force or _stale(op, [sp] + hdrs):
            cmd = [nvcc, *NVCC_FLAGS, "-c", sp, "-o", op]
            if src.endswith(".cc"):
                cmd = [nvcc, *NVCC_FLAGS, "-x", "cu", "-c", sp, "-o", op]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose and (r.stdout or r.stderr):
            print(r.stdout + r.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if force or jobs or _stale(LIB_PATH, objs):
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs])
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in os.sys.argv, verbose=True))
