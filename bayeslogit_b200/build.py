"""Build the engine's shared library in-tree with nvcc for sm_100a.

    python -m bayeslogit_b200.build [--force]

Produces bayeslogit_b200/lib/libbayeslogit_b200.so (git-ignored; it travels to
the GPU box with the working-tree snapshot).  nvcc cross-compiles without a GPU.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libbayeslogit_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-Xptxas", "-v", "--expt-relaxed-constexpr"] + os.environ.get("BL_EXTRA_NVCC_FLAGS", "").split()


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def _compile(src, deps, force, verbose):
    obj = os.path.join(OBJ, os.path.basename(src) + ".o")
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= _newest([src] + deps):
        return obj, ""
    cmd = [NVCC, *ARCH, *CFLAGS, "-I", os.path.join(HERE, "..", "include"), "-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    return obj, p.stderr


def build_native(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, deps, force, verbose), srcs))
    objs = [o for o, _ in res]
    log = "".join(l for _, l in res)
    if verbose and log:
        print(log)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < _newest(objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return LIB


if __name__ == "__main__":
    path = build_native(force="--force" in sys.argv, verbose=True)
    print("built", path)
