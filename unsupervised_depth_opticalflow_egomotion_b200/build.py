"""Build the sm_100a shared library in-tree (``libugl_b200.so`` next to this file).

    python -m unsupervised_depth_opticalflow_egomotion_b200.build [--force] [--verbose]

nvcc cross-compiles on a CPU-only box; the built ``.so`` is git-ignored but travels with the
working tree to the GPU box.  The product has no CPU path: importing the ops without this library
raises.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libugl_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-warn-spills",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libugl_b200.so")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


OBJ_DIR = os.path.join(HERE, "build")


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))


def _compile_one(src: str, obj: str, flags, verbose: bool) -> str:
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "-shared"] + list(flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode:
        raise RuntimeError("nvcc failed on %s (exit %d)\n%s" % (os.path.basename(src), res.returncode, res.stdout + res.stderr))
    return res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False, extra_flags=(), out: str = LIB) -> str:
    """One object per translation unit (compiled in parallel, re-used while neither the source nor any header changed), then a
    link.  ``extra_flags`` / a non-default ``out`` build into their own object directory."""
    if not force and out == LIB and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = OBJ_DIR if (out == LIB and not extra_flags) else out + ".objs"
    os.makedirs(obj_dir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        logs = list(pool.map(lambda j: _compile_one(j[0], j[1], extra_flags, verbose), jobs))
    if verbose:
        sys.stderr.write("".join(logs))
    res = subprocess.run([_nvcc(), "-shared", "-o", out] + objs, capture_output=True, text=True)
    if res.returncode:
        raise RuntimeError("link failed (exit %d)\n%s" % (res.returncode, res.stdout + res.stderr))
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
