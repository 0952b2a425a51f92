"""In-tree build of libctdirect_b200.so with nvcc for sm_100a (no torch extension machinery needed:
the library is a plain C-ABI shared object).

    python -m ctdirect_b200.build [--force] [--only PATTERN] [-j N]
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import fnmatch
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "generated")
OBJ = os.path.join(CSRC, "_build")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libctdirect_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "-diag-suppress", "177", "-Xfatbin", "-compress-all",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    return p


def _common_deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
    deps.append(os.path.normpath(os.path.join(HERE, "..", "include", "ctdirect_b200.h")))
    return deps


def _sources(only=None):
    srcs = [os.path.join(CSRC, "ctd_model.cu"), os.path.join(CSRC, "ctd_capi.cu"), os.path.join(CSRC, "ctd_comm.cu"), os.path.join(CSRC, "ctd_kkt.cu")]
    manifest = os.path.join(GEN, "MANIFEST.txt")
    if not os.path.exists(manifest):
        raise RuntimeError("csrc/generated/MANIFEST.txt missing: run python -m ctdirect_b200.gen.generate")
    for name in open(manifest).read().split():
        if only and not any(fnmatch.fnmatch(name, pat) for pat in only):
            continue
        srcs.append(os.path.join(GEN, name + ".cu"))
    return srcs


def _compile(nvcc, src, obj, extra, verbose):
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout, r.stderr)
    return obj


def build(force: bool = False, only=None, jobs: int | None = None, verbose: bool = False, extra=()) -> str:
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    common = max(os.path.getmtime(p) for p in _common_deps())
    # A snapshot of the tree without the object cache (gpurun ships the library, not csrc/_build): the library on disk is
    # used as it is when it is newer than every source, instead of recompiling ~270 translation units.
    if not force and not only and os.path.exists(LIB) and not any(f.endswith(".o") for f in os.listdir(OBJ)):
        newest = max([common] + [os.path.getmtime(s_) for s_ in _sources(None)] +
                     [os.path.getmtime(s_[:-3] + ".cuh") for s_ in _sources(None) if os.path.exists(s_[:-3] + ".cuh")])
        if os.path.getmtime(LIB) >= newest:
            return LIB
    tasks = []
    objs = []
    for src in _sources(only):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        dep = max(common, os.path.getmtime(src))
        cuh = src[:-3] + ".cuh"
        if os.path.exists(cuh):
            dep = max(dep, os.path.getmtime(cuh))
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < dep:
            tasks.append((src, obj))
    jobs = jobs or max(1, (os.cpu_count() or 2))
    if tasks:
        with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
            futs = [ex.submit(_compile, nvcc, s, o, list(extra), verbose) for s, o in tasks]
            for f in cf.as_completed(futs):
                f.result()
    # relink also when the library on disk was linked from another set of objects (a --only build) or is older than one
    stamp = os.path.join(OBJ, "linked_objects.txt")
    wanted = "\n".join(sorted(os.path.basename(o) for o in objs))
    stale = (not os.path.exists(stamp) or open(stamp).read() != wanted or
             (os.path.exists(LIB) and any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)))
    if tasks or not os.path.exists(LIB) or force or stale:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
               "-o", LIB, *objs, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        with open(stamp, "w") as fh:
            fh.write(wanted)
    return LIB


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--only", action="append")
    ap.add_argument("-j", type=int, default=None)
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args(argv)
    print(build(a.force, a.only, a.j, a.v))
    return 0


if __name__ == "__main__":
    sys.exit(main())
