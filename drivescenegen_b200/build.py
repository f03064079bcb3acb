"""Build recipe for libdsg_b200.so (sm_100a only; nvcc cross-compiles without a GPU).

Usage: ``python -m drivescenegen_b200.build [--force] [--verbose]``.  The library is built IN-TREE next to this file so
that it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libdsg_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    with open(os.path.join(HERE, "..", "include", "dsg_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build_variant(out_path: str, defines, verbose: bool = False) -> str:
    """A/B builds: the same sources with extra -D macros into a separate .so (swapped in on the GPU box by the A/B
    stage of tools/gpu.sh); objects go to their own directory so the default build's stamp stays valid."""
    obj_dir = OBJ + "_variant"
    os.makedirs(obj_dir, exist_ok=True)
    extra = [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(obj_dir, src[:-3] + ".o")
        r = subprocess.run([NVCC, *ARCH, *FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj],
                           capture_output=True, text=True)
        return src, obj, r

    objs = []
    with ThreadPoolExecutor(max_workers=8) as ex:
        for src, obj, r in ex.map(compile_one, _sources()):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- {src}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(obj)
    r = subprocess.run([NVCC, *ARCH, "-shared", "-o", out_path, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return out_path


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [NVCC, *ARCH, *FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with ThreadPoolExecutor(max_workers=8) as ex:
        for src, obj, r in ex.map(compile_one, _sources()):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- {src}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(obj)
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:   # python -m drivescenegen_b200.build --variant out.bin -DNAME=VALUE ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a[2:] for a in sys.argv[i + 2:] if a.startswith("-D")]))
    else:
        print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
