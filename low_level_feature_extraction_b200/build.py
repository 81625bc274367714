"""Build libllfe.so (sm_100a only) in-tree with nvcc.

    python low_level_feature_extraction_b200/build.py [--force]

nvcc cross-compiles without a GPU.  Objects go to csrc/build/, the shared
library next to this file so that it travels with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libllfe.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "llfe.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    common = [nvcc, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "-I", os.path.join(ROOT, "include"), "-I", CSRC]
    if verbose:
        common += ["-Xptxas", "-v"]
    hdr_t = max(os.path.getmtime(p) for p in _deps() if not p.endswith(".cu"))

    def compile_one(src):
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, ""
        r = subprocess.run(common + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log, file=sys.stderr)
    r = subprocess.run([nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
