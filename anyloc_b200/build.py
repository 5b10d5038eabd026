"""Builds libanyloc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).
Every csrc/*.cu is compiled to an object under csrc/build/ (in parallel, only when stale), then linked."""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libanyloc_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def sources():
    """every csrc/*.cu"""
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(os.path.dirname(HERE), "include", "anyloc_b200.h")]


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(LIB, sources() + _headers())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()

    def compile_one(src):
        o = _obj(src)
        if not force and not _stale(o, [src] + hdrs):
            return None
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", o] + (["-Xptxas", "-v"] if verbose else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        logs = list(ex.map(compile_one, sources()))
    if verbose:
        print("\n".join(l for l in logs if l))
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB] +
                       [_obj(s) for s in sources()] + ["-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libanyloc_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
