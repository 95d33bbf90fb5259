"""Build ``libcm2.so`` (the C-ABI kernel library) in-tree with nvcc for sm_100a.

``python -m centermask2_b200.build [--force]``.  nvcc cross-compiles without a GPU; the resulting
``centermask2_b200/libcm2.so`` is git-ignored but travels to the GPU box with the snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libcm2.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "cm2.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for {}:\n{}".format(src, r.stdout + r.stderr))
    return src, r.stderr


def build(force=False, verbose=False):
    """Compile every ``csrc/*.cu`` that is newer than its object and link ``libcm2.so``."""
    os.makedirs(OBJ, exist_ok=True)
    dep = _deps_mtime()
    todo = []
    for s in _sources():
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(dep, os.path.getmtime(os.path.join(CSRC, s))):
            todo.append(s)
    logs = {}
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for src, log in ex.map(lambda s: _compile(s, verbose), todo):
                logs[src] = log
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in _sources()]
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB, logs


if __name__ == "__main__":
    lib, logs = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    for k, v in logs.items():
        if v.strip():
            print("==", k)
            print(v)
    print(lib)
