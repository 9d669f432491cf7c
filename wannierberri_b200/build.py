"""Build libwbgpu.so in-tree for sm_100a:  python -m wannierberri_b200.build [--force] [-v]

Every `csrc/*.cu` is one translation unit, compiled to `csrc/_obj/*.o` (in parallel, only when it or a header it
includes changed) and linked into `libwbgpu.so`."""
import concurrent.futures
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.path.join(HERE, "libwbgpu.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def units():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))] + \
        [os.path.join(os.path.dirname(HERE), "include", "wbgpu.h")]


def _deps(path, seen=None):
    """the file and everything it includes with quotes, recursively"""
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.isfile(path):
        return seen
    seen.add(path)
    for inc in re.findall(r'^\s*#include\s+"([^"]+)"', open(path).read(), flags=re.M):
        _deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    cmd = [os.environ.get("NVCC", "nvcc")] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    todo, objs = [], []
    for src in units():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.isfile(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in _deps(src)):
            todo.append(src)
    if not todo and os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(o) for o in objs):
        return OUT
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as pool:
        list(pool.map(lambda s: _compile(s, verbose), todo))
    cmd = [os.environ.get("NVCC", "nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs
    print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
