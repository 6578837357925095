"""Builds libzkodst.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

Each source is compiled to an object under build/ (only when it or a header changed), then
linked into zk-odst_b200/libzkodst.so.

Usage: python zk-odst_b200/build.py [--force]
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libzkodst.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-Wno-deprecated-gpu-targets", "--extended-lambda",
]


def sources():
    return sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp"))
    )


def headers():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    out.append(os.path.join(HERE, "..", "include", "zkodst.h"))
    return out


def compile_one(src, force):
    obj = os.path.join(OBJ, os.path.basename(src) + ".o")
    newest = max(os.path.getmtime(p) for p in [src] + headers())
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = " ".join(cmd) + "\n" + res.stdout + res.stderr
    if res.returncode != 0:
        raise RuntimeError("nvcc failed on %s\n%s" % (src, log))
    return obj, log


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    logs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(lambda s: compile_one(s, force), srcs))
    objs = [r[0] for r in results]
    logs = [r[1] for r in results if r[1]]
    if logs or not os.path.exists(OUT):
        cmd = [NVCC, "--shared", "-o", OUT] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("link failed\n" + logs[-1])
    if logs:
        with open(os.path.join(HERE, "build.log"), "a" if not force else "w") as f:
            f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
