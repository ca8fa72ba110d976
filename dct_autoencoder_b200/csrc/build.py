"""Builds libdcta.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python dct_autoencoder_b200/csrc/build.py [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libdcta.so")
STAMP = os.path.join(PKG, ".libdcta.stamp")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
    "--fmad=true",
]


def sources():
    return sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cu"))


def digest():
    h = hashlib.sha256()
    root = os.path.dirname(PKG)
    files = sources() + [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith((".cuh", ".h"))]
    files.append(os.path.join(root, "include", "dcta.h"))
    for f in files:
        h.update(f.encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _headers():
    root = os.path.dirname(PKG)
    hs = [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith((".cuh", ".h"))]
    return hs + [os.path.join(root, "include", "dcta.h")]


def _obj_digest(src):
    h = hashlib.sha256()
    for f in [src] + _headers():
        h.update(open(f, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()[:24]


def build(force=False, verbose=False):
    """One object per source, compiled in parallel and cached by content under csrc/_obj (git-ignored, not sent to
    the GPU box), then linked into libdcta.so."""
    d = digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read().strip() == d:
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in FLAGS if f != "--shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + "." + _obj_digest(src) + ".o")
        if force or not os.path.exists(obj):
            for old in os.listdir(objdir):
                if old.startswith(os.path.basename(src)[:-3] + "."):
                    os.remove(os.path.join(objdir, old))
            cmd = [NVCC] + cflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            print("[dcta build]", " ".join(cmd), flush=True)
            subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [NVCC, "--shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", OUT, "-lcudart"]
    print("[dcta build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(d)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)
